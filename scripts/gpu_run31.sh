mkdir -p gpurun_out
timeout -k 10 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout -k 10 600 python __graft_entry__.py smoke 2>&1 | tail -2
timeout -k 10 900 python bench.py > gpurun_out/r31_default.json 2> gpurun_out/r31_default.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r31_default.json; echo; tail -n 2 gpurun_out/r31_default.err
timeout -k 10 900 python bench.py --impl reference > gpurun_out/r31_reference.json 2> gpurun_out/r31_reference.err; echo "ref rc=$?"
tail -c 900 gpurun_out/r31_reference.json; echo; tail -n 2 gpurun_out/r31_reference.err
