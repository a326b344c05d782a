mkdir -p gpurun_out
( timeout -k 10 400 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_sharded.py tests/test_bench_contract.py -q -m gpu -x -s 2>&1 | grep -v "^frame" | tail -n 14 ) > gpurun_out/r16_tests.log 2>&1
( timeout -k 10 600 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2_hybrid_n1.json 2> gpurun_out/r2_hybrid_n1.err
cat gpurun_out/r16_tests.log | cut -c1-300
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_hybrid_n1.json").read().strip().splitlines()[-1])
print(round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "pipe", round(d["e2e_pipeline"]["value"]), {k:v for k,v in d["parity"].items() if k in ("queries_with_identical_topk","verify_flagged_queries_in_timed_steps","sharded_digest_equal")})
for k in d["roofline"]["kernels"]: print("  ", k["name"][:50], round(k["ms_per_step"],3), round(k["frac_hbm"],3))
for p in d.get("points") or []: print("  pt", p["queries_per_step"], p["dense_mode"], round(p["value"]), round(p["ms_per_step"],3))
print(d["clocks"], d["roofline"]["frac"], d["roofline"]["kernel"][:40])
PY
tail -n 3 gpurun_out/r2_hybrid_n1.err
