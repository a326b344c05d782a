mkdir -p gpurun_out
timeout -k 10 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r42_tests.log 2>&1; echo "rc=$?"
grep -n "Error\|error\|FAILED\|passed\|failed\|test_" gpurun_out/r42_tests.log | head -30
