mkdir -p gpurun_out
( timeout -k 10 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_sharded.py -q -m gpu -x 2>&1 | tail -n 4 ) > gpurun_out/r17_tests.log 2>&1
CMD="python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline"
$CMD > gpurun_out/r17_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"dense_gemm|gemm_prepare|bm25_|fuse_|topk_|verify_|keys_|stats_|cand_" --csv --log-file gpurun_out/r2_launches_bench_default.csv $CMD > gpurun_out/r17_ncu1.log 2>&1
python scripts/profile_gemm.py hybrid_bf16x 10000000 > gpurun_out/r17_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dense_gemm_kernel -s 1 -c 1 -o gpurun_out/prof_r2_gemm_bf16_ext python scripts/profile_gemm.py hybrid_bf16x 10000000 > gpurun_out/r17_ncu2.log 2>&1
cat gpurun_out/r17_tests.log; tail -n 2 gpurun_out/r17_ncu1.log gpurun_out/r17_ncu2.log; wc -l gpurun_out/r2_launches_bench_default.csv
