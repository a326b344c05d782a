"""CPU simulation (numpy) of the error of a 3xTF32 split dot product -- the candidate for an fp32-accurate dense
scan on the tensor cores at large query batch.  Operands are truncated to TF32 the way
`tcgen05.mma.kind::tf32` reads 32-bit containers (low 13 mantissa bits ignored), a = a_hi + a_lo with
a_lo = tf32(a - a_hi), products accumulated in float32: hi*hi + hi*lo + lo*hi.

    python scripts/sim_tf32x3.py        ->  max |cos error|: fp32 3.9e-08, 1xTF32 1.8e-04, 3xTF32 1.2e-07
"""
import numpy as np

rng = np.random.default_rng(0)
N, d = 200_000, 384
V = rng.standard_normal((N, d)).astype(np.float32)
q = rng.standard_normal(d).astype(np.float32)


def tf32(x):
    return (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


exact = V.astype(np.float64) @ q.astype(np.float64)
scale = np.linalg.norm(q.astype(np.float64)) * np.linalg.norm(V.astype(np.float64), axis=1)
Vh, qh = tf32(V), tf32(q)
Vl, ql = tf32(V - Vh), tf32(q - qh)
variants = {"fp32": V @ q, "1xTF32": Vh @ qh, "3xTF32": Vh @ qh + Vh @ ql + Vl @ qh,
            "4xTF32": Vh @ qh + Vh @ ql + Vl @ qh + Vl @ ql}
want = np.argsort(-(exact / scale), kind="stable")[:100]
for name, x in variants.items():
    e = np.abs(x.astype(np.float64) - exact) / scale
    same = float(np.mean(np.argsort(-(x / scale), kind="stable")[:100] == want))
    print(f"{name:7s} max |cos error| {e.max():.2e}  mean {e.mean():.2e}  top-100 ids equal to float64: {same:.2f}")
