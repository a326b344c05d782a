mkdir -p gpurun_out
timeout -k 10 600 python scripts/profile_gemm.py hybrid_bf16x256 10000000 > gpurun_out/r22_plain.log 2>&1; echo "plain rc=$?"
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:dense_gemm_kernel -s 1 -c 1 -f -o gpurun_out/prof_r2_gemm_bf16_ext_q256 python scripts/profile_gemm.py hybrid_bf16x256 10000000 > gpurun_out/r22_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r22_ncu.log
