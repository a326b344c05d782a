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


import pytest


@pytest.mark.gpu
def test_bench_runs_on_a_small_corpus_and_reproduces_the_committed_digest():
    """bench.py end to end on the GPU at a size that takes seconds: one JSON line with every contract key, the
    per-kernel rooflines, the plugin-API e2e leg, ids identical to the exact mode, and the exact-mode digest of the fixed
    64-query batch equal to the constant in tests/golden/bench_digests.json (generated at N = 1)."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--n-docs", "300000", "--vocab", "50000",
                          "--batch", "16", "--steps", "3", "--warmup", "3", "--no-cpu-baseline"],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert (BASE_KEYS | {"gpu_launches", "clocks", "roofline", "parity", "e2e_pipeline"}) <= set(d)
    assert d["value"] > 0 and d["e2e"]["value"] > 0 and d["e2e_pipeline"]["value"] > 0 and d["gpu_launches"] > 0
    assert d["e2e"]["same_ids_as_device_run"] and d["e2e_pipeline"]["same_ids_as_device_run"]
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["parity"]["sharded_digest_equal"] is True
    assert d["parity"]["queries_with_identical_topk"] >= 0.9
    names = [k["name"] for k in d["roofline"]["kernels"]]
    assert len(names) == 4 and names[-1].startswith("step")
    assert all(k["ms_per_step"] > 0 and k["alg_bytes_per_step"] > 0 for k in d["roofline"]["kernels"])
