mkdir -p gpurun_out
timeout -k 10 900 python bench.py --workload bm25 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r33_bm25_n1.json 2> gpurun_out/r33_bm25_n1.err; echo "bm25 rc=$?"
timeout -k 10 900 python bench.py --workload multi_stage --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r33_multi_stage_n1.json 2> gpurun_out/r33_multi_stage_n1.err; echo "ms rc=$?"
timeout -k 10 900 python bench.py --workload diversity --steps 3 --warmup 3 --batch 4096 --no-cpu-baseline > gpurun_out/r33_diversity_n1.json 2> gpurun_out/r33_diversity_n1.err; echo "div rc=$?"
timeout -k 10 900 python bench.py --steps 20 --warmup 5 --batch 8 --dense-mode fp32 --no-extras --no-cpu-baseline > gpurun_out/r33_hybrid_b8_n1.json 2> gpurun_out/r33_hybrid_b8_n1.err; echo "b8 rc=$?"
python - <<PY
import json
for f in ('bm25','multi_stage','diversity','hybrid_b8'):
    d=json.loads(open(f'gpurun_out/r33_{f}_n1.json').read().strip().splitlines()[-1])
    print(f, round(d['value']), round(d['ms_per_step'],3), 'e2e', d.get('e2e',{}).get('value'), 'roof', d['roofline'].get('frac'))
PY
