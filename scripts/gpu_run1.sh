mkdir -p gpurun_out
( timeout -k 10 200 python -m pytest tests/test_gpu_gemm.py -q -m gpu -k "bf16_two_query or doc_range" 2>&1 | tail -40 ) > gpurun_out/r2_gemm_a.log 2>&1
( timeout -k 10 200 python -m pytest tests/test_gpu_gemm.py -q -m gpu -k "tf32x3_within" 2>&1 | tail -60 ) > gpurun_out/r2_gemm_b.log 2>&1
( timeout -k 10 300 python -m pytest tests/test_gpu_gemm.py -q -m gpu -k "filtered or hybrid_ids" -s 2>&1 | tail -60 ) > gpurun_out/r2_gemm_c.log 2>&1
( timeout -k 10 600 python -m pytest tests/test_gpu_round2.py -q -m gpu -s 2>&1 | tail -60 ) > gpurun_out/r2_round2.log 2>&1
( timeout -k 10 600 python -m pytest tests -q -m gpu --deselect tests/test_gpu_gemm.py --deselect tests/test_gpu_round2.py 2>&1 | tail -40 ) > gpurun_out/r2_rest.log 2>&1
tail -5 gpurun_out/r2_gemm_a.log gpurun_out/r2_gemm_b.log gpurun_out/r2_gemm_c.log gpurun_out/r2_round2.log gpurun_out/r2_rest.log
