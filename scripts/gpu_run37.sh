mkdir -p gpurun_out
timeout -k 10 600 python scripts/profile_gemm.py filter_bf16 10000000 768 > gpurun_out/r37_plain.log 2>&1; echo "plain rc=$?"
timeout -k 10 1200 ncu --set full --clock-control none --import-source on -k regex:"dense_gemm_kernel" -s 3 -c 1 -f -o gpurun_out/prof_r2_final_filter_d768 python scripts/profile_gemm.py filter_bf16 10000000 768 > gpurun_out/r37_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r37_ncu.log
