mkdir -p gpurun_out
( timeout -k 10 400 python -m pytest tests/test_gpu_gemm.py -q -m gpu -x -s -k "bf16_exact" 2>&1 | grep -v "^frame" | tail -n 30 ) > gpurun_out/r13_tests.log 2>&1
( timeout -k 10 300 python bench.py --dense-mode bf16_exact --batch 256 --steps 10 --warmup 3 --no-extras --no-cpu-baseline ) > gpurun_out/r13_bench_bf16x.json 2> gpurun_out/r13_bench_bf16x.err
( timeout -k 10 300 python bench.py --dense-mode bf16_exact --batch 128 --steps 10 --warmup 3 --no-extras --no-cpu-baseline ) > gpurun_out/r13_bench_bf16x_128.json 2>> gpurun_out/r13_bench_bf16x.err
cat gpurun_out/r13_tests.log | cut -c1-300; tail -c 1800 gpurun_out/r13_bench_bf16x.json; echo; tail -c 600 gpurun_out/r13_bench_bf16x_128.json; tail -n 5 gpurun_out/r13_bench_bf16x.err
