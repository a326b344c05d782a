# usage: bash scripts/gpu_scale2.sh N tag  -- hybrid default + multi_stage at N ranks
N=$1; TAG=${2:-r2}
mkdir -p gpurun_out
run() {
  name=$1; shift
  if [ "$N" = "1" ]; then
    timeout -k 10 900 python bench.py --gpus 1 "$@" > gpurun_out/${TAG}_${name}_n1.json 2> gpurun_out/${TAG}_${name}_n1.err
  else
    timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N "$@" > gpurun_out/${TAG}_${name}_n$N.json 2> gpurun_out/${TAG}_${name}_n$N.err
  fi
  echo "== $name n=$N rc=$?"; tail -c 300 gpurun_out/${TAG}_${name}_n$N.json; echo; tail -n 2 gpurun_out/${TAG}_${name}_n$N.err | cut -c1-300
}
run hybrid --steps 20 --warmup 5 --no-cpu-baseline
run multi_stage --workload multi_stage --steps 5 --warmup 3 --no-cpu-baseline
