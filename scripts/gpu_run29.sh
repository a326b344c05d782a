mkdir -p gpurun_out
for v in 0 1; do
HS_SCREEN_BM25_F32=$v timeout -k 10 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r29_v$v.json 2> gpurun_out/r29_v$v.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r29_v$v.json').read().strip().splitlines()[-1])
print('bm25 f32 screen=$v: q/s', round(d['value']), 'step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), [round(k['ms_per_step'],3) for k in d['roofline']['kernels']], d['parity'].get('verify_flagged_queries_in_timed_steps'), d['parity'].get('sharded_digest_equal'))
print(d['roofline']['kernels'][2].get('parts_ms_rank0'))
PY
done
