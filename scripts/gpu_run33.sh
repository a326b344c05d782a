mkdir -p gpurun_out
timeout -k 10 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for b in 256 8; do
extra=""; if [ $b = 8 ]; then extra="--dense-mode fp32"; fi
timeout -k 10 600 python bench.py --steps 20 --warmup 5 --batch $b $extra --no-extras --no-cpu-baseline > gpurun_out/r33_b$b.json 2> gpurun_out/r33_b$b.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r33_b$b.json').read().strip().splitlines()[-1])
print('B=$b q/s', round(d['value']), 'step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), [round(k['ms_per_step'],3) for k in d['roofline']['kernels']], d['parity'].get('verify_flagged_queries_in_timed_steps'), d['parity'].get('sharded_digest_equal'))
PY
done
timeout -k 10 600 python bench.py --workload bm25 --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r33_bm25.json 2> gpurun_out/r33_bm25.err; python -c "
import json
d=json.loads(open('gpurun_out/r33_bm25.json').read().strip().splitlines()[-1]); print('bm25 cfg3', d['value'], d['ms_per_step'], [round(k['ms_per_step'],3) for k in d['roofline']['kernels']])"
