mkdir -p gpurun_out
( timeout -k 10 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_fullsize.py tests/test_bench_contract.py -q -m gpu -x 2>&1 | tail -15 ) > gpurun_out/r4_tests.log 2>&1
( HS_BM25_HOT=0 timeout -k 10 300 python scripts/bench_gemm.py --which hybrid ) > gpurun_out/r4_hybrid_nohot.jsonl 2> gpurun_out/r4_hybrid_nohot.err
( timeout -k 10 300 python scripts/bench_gemm.py --which hybrid ) > gpurun_out/r4_hybrid_hot.jsonl 2> gpurun_out/r4_hybrid_hot.err
python scripts/profile_gemm.py filter_bf16 4000000 > gpurun_out/r4_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r4_launches_filter_bf16.csv python scripts/profile_gemm.py filter_bf16 4000000 > gpurun_out/r4_ncu.log 2>&1
tail -n 4 gpurun_out/r4_tests.log; cat gpurun_out/r4_hybrid_nohot.jsonl gpurun_out/r4_hybrid_hot.jsonl; tail -n 3 gpurun_out/r4_ncu.log
