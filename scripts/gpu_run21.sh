mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_round2.py -m gpu -x -q 2>&1 | tail -8
timeout -k 10 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r21_default.json 2> gpurun_out/r21_default.err; echo "rc=$?"
tail -c 300 gpurun_out/r21_default.json; tail -n 3 gpurun_out/r21_default.err
timeout -k 10 600 python scripts/bench_gemm.py --which tf32,bf16 > gpurun_out/r21_gemm.jsonl 2> gpurun_out/r21_gemm.err; echo "rc=$?"
cat gpurun_out/r21_gemm.jsonl | cut -c1-330
