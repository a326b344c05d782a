# verify_topk with 32 warps, default batch 256, chain split
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_gemm.py tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -5
timeout -k 10 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r18_default.json 2> gpurun_out/r18_default.err; echo "rc=$?"
tail -c 1500 gpurun_out/r18_default.json; tail -n 3 gpurun_out/r18_default.err
