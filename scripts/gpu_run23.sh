mkdir -p gpurun_out
for st in 0 4 3 2; do
HS_GEMM_MAX_STAGES=$st timeout -k 10 600 python bench.py --steps 10 --warmup 3 --batch 128 --no-extras --no-cpu-baseline > gpurun_out/r23_b128_s$st.json 2> gpurun_out/r23_b128_s$st.err; echo "stages $st rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r23_b128_s$st.json').read().strip().splitlines()[-1])
print('stages cap $st: step', round(d['ms_per_step'],3), 'gemm', round(d['roofline']['kernels'][0]['ms_per_step'],3))
PY
done
