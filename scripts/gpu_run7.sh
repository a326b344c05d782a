mkdir -p gpurun_out
( timeout -k 10 240 python -m pytest tests/test_gpu_gemm.py -q -m gpu -x 2>&1 | grep -v "^frame" | tail -n 25 ) > gpurun_out/r7_gemm_tests.log 2>&1
if grep -q "passed" gpurun_out/r7_gemm_tests.log && ! grep -q "failed\|Timeout\|Killed" gpurun_out/r7_gemm_tests.log; then
  for cl in 1 2 4; do
    ( HS_GEMM_CLUSTER=$cl timeout -k 10 300 python scripts/bench_gemm.py --which filter,cfg4 ) > gpurun_out/r7_gemm_cl$cl.jsonl 2> gpurun_out/r7_gemm_cl$cl.err
  done
fi
( timeout -k 10 200 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -q -m gpu -k "mmr or diversity" 2>&1 | tail -n 5 ) > gpurun_out/r7_mmr_tests.log 2>&1
( timeout -k 10 300 python scripts/bench_configs.py --which 5 ) > gpurun_out/r7_mmr.jsonl 2> gpurun_out/r7_mmr.err
cat gpurun_out/r7_gemm_tests.log; for cl in 1 2 4; do echo "== cluster $cl"; cat gpurun_out/r7_gemm_cl$cl.jsonl; tail -n 2 gpurun_out/r7_gemm_cl$cl.err; done; cat gpurun_out/r7_mmr_tests.log gpurun_out/r7_mmr.jsonl
