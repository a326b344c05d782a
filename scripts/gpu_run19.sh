mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_parity.py -m gpu -x -q -k "bf16_exact or serving" 2>&1 | tail -5
timeout -k 10 300 python scripts/profile_host.py 256 > gpurun_out/r19_host.txt 2>&1; echo "rc=$?"
head -60 gpurun_out/r19_host.txt | cut -c1-170
timeout -k 10 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r19_default.json 2> gpurun_out/r19_default.err; echo "rc=$?"
tail -c 300 gpurun_out/r19_default.json; tail -n 3 gpurun_out/r19_default.err
