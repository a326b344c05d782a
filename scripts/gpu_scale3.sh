N=$1; TAG=${2:-r2}
mkdir -p gpurun_out
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_hybrid_n$N.json 2> gpurun_out/${TAG}_hybrid_n$N.err
echo "rc=$?"; tail -c 200 gpurun_out/${TAG}_hybrid_n$N.json
