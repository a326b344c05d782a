mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -q -m gpu -x -p no:cacheprovider > gpurun_out/r15_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r15_tests.log
( timeout -k 10 600 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2_hybrid_n1.json 2> gpurun_out/r2_hybrid_n1.err
grep -v "^frame\|^$" gpurun_out/r15_tests.log | tail -n 12 | cut -c1-300
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_hybrid_n1.json").read().strip().splitlines()[-1])
print(round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "pipe", round(d["e2e_pipeline"]["value"]), {k:v for k,v in d["parity"].items() if k in ("queries_with_identical_topk","verify_flagged_queries_in_timed_steps","sharded_digest_equal")})
for k in d["roofline"]["kernels"]: print("  ", k["name"][:50], round(k["ms_per_step"],3), round(k["frac_hbm"],3))
for p in d.get("points") or []: print("  pt", p["queries_per_step"], p["dense_mode"], round(p["value"]), round(p["ms_per_step"],3))
print(d["clocks"])
PY
tail -n 3 gpurun_out/r2_hybrid_n1.err
