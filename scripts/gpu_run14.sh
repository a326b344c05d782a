mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -q -m gpu -x -p no:cacheprovider > gpurun_out/r14_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r14_tests.log
for b in 128 256; do
( timeout -k 10 300 python bench.py --dense-mode bf16_exact --batch $b --steps 10 --warmup 3 --no-extras --no-cpu-baseline ) > gpurun_out/r14_bench_bf16x_$b.json 2> gpurun_out/r14_bench_bf16x_$b.err
done
( timeout -k 10 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline ) > gpurun_out/r14_bench_tf32.json 2> gpurun_out/r14_bench_tf32.err
( timeout -k 10 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline --batch 8 --dense-mode fp32 ) > gpurun_out/r14_bench_b8.json 2> gpurun_out/r14_bench_b8.err
grep -v "^frame\|^$" gpurun_out/r14_tests.log | tail -n 12 | cut -c1-300
python - <<'PY'
import json
for f in ["bf16x_128","bf16x_256","tf32","b8"]:
    try:
        d=json.loads(open(f"gpurun_out/r14_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), {k:v for k,v in d["parity"].items() if k in ("queries_with_identical_topk","verify_flagged_queries_in_timed_steps","sharded_digest_equal")}, [(k["name"][:12], round(k["ms_per_step"],3)) for k in d["roofline"]["kernels"]])
    except Exception as e:
        print(f, "ERR", e, open(f"gpurun_out/r14_bench_{f}.err").read()[-800:])
PY
