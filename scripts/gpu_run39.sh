mkdir -p gpurun_out
for m in 128 256 512 1024; do
HS_TOPK_LIST_MULT=$m timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r39_m$m.json 2> gpurun_out/r39_m$m.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r39_m$m.json').read().strip().splitlines()[-1])
print('mult=$m N=2: q/s', round(d['value']), 'step', round(d['ms_per_step'],3), [round(k['ms_per_step'],3) for k in d['roofline']['kernels']], d['roofline']['kernels'][2].get('parts_ms_rank0'), d['parity'].get('sharded_digest_equal'))
PY
done
