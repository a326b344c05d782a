mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -q -m gpu -x -p no:cacheprovider > gpurun_out/r5_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r5_tests.log
( timeout -k 10 400 python scripts/bench_gemm.py --which filter,cfg4 ) > gpurun_out/r5_gemm_bench.jsonl 2> gpurun_out/r5_gemm_bench.err
( timeout -k 10 300 python scripts/bench_configs.py --which 5 ) > gpurun_out/r5_mmr_cluster.jsonl 2> gpurun_out/r5_mmr.err
( HS_MMR_NO_CLUSTER=1 timeout -k 10 300 python scripts/bench_configs.py --which 5 ) > gpurun_out/r5_mmr_nocluster.jsonl 2>> gpurun_out/r5_mmr.err
grep -v "^frame\|^$" gpurun_out/r5_tests.log | tail -n 40 | cut -c1-400; cat gpurun_out/r5_gemm_bench.jsonl gpurun_out/r5_mmr_cluster.jsonl gpurun_out/r5_mmr_nocluster.jsonl; tail -n 3 gpurun_out/r5_gemm_bench.err gpurun_out/r5_mmr.err
