mkdir -p gpurun_out
python scripts/profile_gemm.py hybrid_tf32 10000000 > gpurun_out/r12_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:bm25_batch_kernel -s 1 -c 1 -o gpurun_out/prof_r2_bm25_batch_hot python scripts/profile_gemm.py hybrid_tf32 10000000 > gpurun_out/r12_ncu.log 2>&1
tail -n 3 gpurun_out/r12_ncu.log
