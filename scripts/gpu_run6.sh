mkdir -p gpurun_out
( timeout -k 10 200 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -q -m gpu -k "mmr or diversity" 2>&1 | tail -n 5 ) > gpurun_out/r6_mmr_tests.log 2>&1
( timeout -k 10 300 python scripts/bench_configs.py --which 5 ) > gpurun_out/r6_mmr_cluster.jsonl 2> gpurun_out/r6_mmr.err
python scripts/profile_gemm.py hybrid_tf32 10000000 > gpurun_out/r6_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dense_gemm_kernel -s 1 -c 1 -o gpurun_out/prof_r2_gemm_tf32x3 python scripts/profile_gemm.py hybrid_tf32 10000000 > gpurun_out/r6_ncu.log 2>&1
cat gpurun_out/r6_mmr_tests.log gpurun_out/r6_mmr_cluster.jsonl; tail -n 3 gpurun_out/r6_mmr.err gpurun_out/r6_ncu.log
