mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_round2.py tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -8
timeout -k 10 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r20_default.json 2> gpurun_out/r20_default.err; echo "rc=$?"
tail -c 300 gpurun_out/r20_default.json; tail -n 3 gpurun_out/r20_default.err
HS_SCREEN_F32=1 timeout -k 10 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r20_f32screen.json 2> gpurun_out/r20_f32screen.err; echo "rc=$?"
timeout -k 10 600 python bench.py --steps 20 --warmup 5 --batch 128 --no-extras --no-cpu-baseline > gpurun_out/r20_b128.json 2> gpurun_out/r20_b128.err; echo "rc=$?"
