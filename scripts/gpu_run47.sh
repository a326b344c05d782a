mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r47_tests.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r47_tests.log | cut -c1-200
timeout -k 10 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout -k 10 900 python bench.py > gpurun_out/r2f_hybrid_n1.json 2> gpurun_out/r47_default.err; echo "bench rc=$?"
