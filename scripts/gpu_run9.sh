mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -q -m gpu -x -p no:cacheprovider > gpurun_out/r9_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r9_tests.log
( timeout -k 10 500 python scripts/bench_gemm.py --which tf32,bf16,filter,cfg4 ) > gpurun_out/r9_gemm_bench.jsonl 2> gpurun_out/r9_gemm_bench.err
grep -v "^frame\|^$" gpurun_out/r9_tests.log | tail -n 30 | cut -c1-300; cat gpurun_out/r9_gemm_bench.jsonl; tail -n 3 gpurun_out/r9_gemm_bench.err
