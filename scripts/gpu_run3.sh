mkdir -p gpurun_out
( timeout -k 10 300 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_gemm.py -q -m gpu -x 2>&1 | tail -30 ) > gpurun_out/r3_tests.log 2>&1
( timeout -k 10 200 python bench.py --n-docs 300000 --vocab 50000 --batch 16 --steps 3 --warmup 3 --no-cpu-baseline --dense-mode tf32x3 ) > gpurun_out/r3_bench_small.json 2> gpurun_out/r3_bench_small.err
( timeout -k 10 600 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r3_bench_default.json 2> gpurun_out/r3_bench_default.err
( timeout -k 10 300 python bench.py --steps 20 --warmup 5 --batch 8 --dense-mode fp32 --no-extras --no-cpu-baseline ) > gpurun_out/r3_bench_b8.json 2> gpurun_out/r3_bench_b8.err
( timeout -k 10 400 python scripts/bench_gemm.py --which filter,cfg4 ) > gpurun_out/r3_gemm_bench.jsonl 2> gpurun_out/r3_gemm_bench.err
tail -3 gpurun_out/r3_tests.log; tail -2 gpurun_out/*.err; wc -c gpurun_out/r3_*.json*
