mkdir -p gpurun_out
( timeout -k 10 300 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_round2.py -q -m gpu -x 2>&1 | grep -v "^frame" | tail -n 12 ) > gpurun_out/r11_tests.log 2>&1
( timeout -k 10 500 python scripts/bench_gemm.py --which tf32,bf16,filter,cfg4 ) > gpurun_out/r11_gemm_bench.jsonl 2> gpurun_out/r11_gemm_bench.err
python scripts/sanitize_small.py > gpurun_out/r11_sanitize_plain.log 2>&1 && timeout -k 10 900 compute-sanitizer --tool memcheck --error-exitcode 7 python scripts/sanitize_small.py > gpurun_out/r11_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r11_memcheck.log
cat gpurun_out/r11_tests.log gpurun_out/r11_gemm_bench.jsonl; tail -n 3 gpurun_out/r11_gemm_bench.err; tail -n 12 gpurun_out/r11_memcheck.log
