mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_round2.py tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -8
for b in 256 128; do
timeout -k 10 600 python bench.py --steps 20 --warmup 5 --batch $b --no-extras --no-cpu-baseline > gpurun_out/r28_b$b.json 2> gpurun_out/r28_b$b.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r28_b$b.json').read().strip().splitlines()[-1])
print('B=$b: q/s', round(d['value']), 'step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), [round(k['ms_per_step'],3) for k in d['roofline']['kernels']], d['parity'].get('verify_flagged_queries_in_timed_steps'), d['parity'].get('sharded_digest_equal'))
print(d['roofline']['kernels'][2].get('parts_ms_rank0'))
PY
done
