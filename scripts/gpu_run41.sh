mkdir -p gpurun_out
timeout -k 10 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout -k 10 600 python __graft_entry__.py smoke 2>&1 | tail -2
timeout -k 10 900 python bench.py > gpurun_out/r2f_hybrid_n1.json 2> gpurun_out/r41_default.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2f_hybrid_n1.json; echo; tail -n 2 gpurun_out/r41_default.err
