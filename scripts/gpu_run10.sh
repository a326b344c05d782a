mkdir -p gpurun_out
python scripts/profile_gemm.py filter_bf16 4000000 > gpurun_out/r10_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dense_gemm_kernel -s 3 -c 1 -o gpurun_out/prof_r2_gemm_bf16_filter_v2 python scripts/profile_gemm.py filter_bf16 4000000 > gpurun_out/r10_ncu.log 2>&1
tail -n 3 gpurun_out/r10_ncu.log
