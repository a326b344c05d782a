mkdir -p gpurun_out
timeout -k 10 600 python scripts/profile_gemm.py hybrid_bf16x256 10000000 > gpurun_out/r32_plain.log 2>&1; echo "plain rc=$?"
timeout -k 10 1200 ncu --set full --clock-control none --import-source on -k regex:"bm25_batch_kernel|dense_gemm_kernel|fuse_topk_kernel" -s 3 -c 3 -f -o gpurun_out/prof_r2_final_b256 python scripts/profile_gemm.py hybrid_bf16x256 10000000 > gpurun_out/r32_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r32_ncu.log
