mkdir -p gpurun_out
for pf in 2 4 8 1; do
HS_BM25_PF=$pf timeout -k 10 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r36_pf$pf.json 2> gpurun_out/r36_pf$pf.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r36_pf$pf.json').read().strip().splitlines()[-1])
print('pf=$pf: q/s', round(d['value']), 'step', round(d['ms_per_step'],3), [round(k['ms_per_step'],3) for k in d['roofline']['kernels']], d['parity'].get('sharded_digest_equal'))
PY
done
