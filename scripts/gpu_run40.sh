mkdir -p gpurun_out
for cfg in "128 2" "512 2" "512 8" "512 16" "256 16"; do
set -- $cfg
HS_TOPK_LIST_MULT=$1 HS_TOPK_FILL=$2 timeout -k 10 600 python bench.py --n-docs 1250000 --steps 30 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r40.json 2> gpurun_out/r40.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r40.json').read().strip().splitlines()[-1])
print('mult=$1 fill=$2 n=1.25M: q/s', round(d['value']), 'step', round(d['ms_per_step'],3), [round(k['ms_per_step'],3) for k in d['roofline']['kernels']], d['roofline']['kernels'][2].get('parts_ms_rank0',{}).get('select'))
PY
done
for cfg in "128 2" "512 16"; do
set -- $cfg
HS_TOPK_LIST_MULT=$1 HS_TOPK_FILL=$2 timeout -k 10 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r40.json 2> gpurun_out/r40.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r40.json').read().strip().splitlines()[-1])
print('mult=$1 fill=$2 n=10M: q/s', round(d['value']), 'step', round(d['ms_per_step'],3), [round(k['ms_per_step'],3) for k in d['roofline']['kernels']], d['roofline']['kernels'][2].get('parts_ms_rank0',{}).get('select'))
PY
HS_TOPK_LIST_MULT=$1 HS_TOPK_FILL=$2 timeout -k 10 600 python bench.py --steps 20 --warmup 5 --batch 8 --dense-mode fp32 --no-extras --no-cpu-baseline > gpurun_out/r40.json 2> gpurun_out/r40.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r40.json').read().strip().splitlines()[-1])
print('mult=$1 fill=$2 n=10M B=8: q/s', round(d['value']), 'step', round(d['ms_per_step'],3), [round(k['ms_per_step'],3) for k in d['roofline']['kernels']])
PY
done
