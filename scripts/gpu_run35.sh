mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline"
$CMD > gpurun_out/r35_plain.json 2> gpurun_out/r35_plain.err && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"dense_gemm|gemm_prepare|bm25_|fuse_|topk_|verify_|keys_|stats_|cand_" --csv --log-file gpurun_out/r2_launches_bench_default_final.csv $CMD > gpurun_out/r35_ncu.log 2>&1
echo "rc=$?"; wc -l gpurun_out/r2_launches_bench_default_final.csv; tail -c 300 gpurun_out/r35_plain.json
