mkdir -p gpurun_out
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2f_hybrid_latest_n2.json 2> gpurun_out/r48.err
echo "rc=$?"; tail -c 200 gpurun_out/r2f_hybrid_latest_n2.json
