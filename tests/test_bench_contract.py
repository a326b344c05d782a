"""bench.py contract on CPU: the reference arm (the UNMODIFIED reference pipeline -- /root/reference here, the staged
copy under oracle/_ref on a GPU box -- timed on the host cores with forked workers; the oracle port when neither is
present) prints exactly one JSON line carrying the keys the driver reads; the CUDA arm's line is checked for the same key set statically (it needs
a GPU to run)."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                          "--warmup", "1", "--ref-docs", "2000", "--ref-workers", "2", "--cpu-sample-docs", "20000"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"].startswith("hybrid_bm25 queries/sec @10M docs") and d["unit"] == "queries/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["n_gpus"] == 1
    assert "workload" in d["config"] and "model" not in d["config"]
    from oracle import refload
    assert d["cpu_baseline"]["kind"] == ("reference" if refload.available() else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    if refload.available():
        assert d["cpu_baseline"]["cores"] == 2 and "unmodified reference" in d["cpu_baseline"]["sample"]
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_cuda_arm_line_has_the_contract_keys():
    src = open(os.path.join(ROOT, "bench.py")).read()
    body = src[src.index("def main_cuda") if "def main_cuda" in src else 0:]
    for key in BASE_KEYS | {"gpu_launches", "clocks", "roofline", "parity"}:
        assert re.search(r'"%s"\s*:' % re.escape(key), body), key
    for key in ("bound", "achieved", "peak", "frac", "traffic", "sm_mhz", "sm_max_mhz", "reasons", "h2d_bytes_per_step",
                "d2h_bytes_per_step", "cores", "kind", "sample", "workload"):
        assert re.search(r'"%s"' % key, src), key
