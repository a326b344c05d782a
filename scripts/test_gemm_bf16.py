"""Bring-up check of the tcgen05 bf16 GEMM dense path against torch (bf16 inputs, fp32 accumulate)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hybrid_search_engine_b200 as hs
from hybrid_search_engine_b200.engine import SearchEngine

torch.manual_seed(0)
ok = True
for n, d, B in [(1000, 384, 5), (128, 64, 16), (5000, 384, 128), (100_000, 384, 200), (3000, 640, 64), (2500, 768, 100), (1500, 1024, 7), (777, 100, 33), (1001, 384, 130)]:
    v = torch.randn(n, d, device="cuda"); q = torch.randn(B, d, device="cuda")
    v[3] = 0; q[min(2, B - 1)] = 0
    shard = hs.DeviceIndex("cuda:0", n); shard.set_dense(v)
    eng = SearchEngine(shard)
    stats = eng._stats(B)
    qd = eng.upload_vectors(q.cpu().numpy()).clone()
    cos = eng.dense_scan(qd, stats, "bf16").clone()
    torch.cuda.synchronize()
    dot = (q.to(torch.bfloat16).float() @ v.to(torch.bfloat16).float().T)
    vn = shard.vnorm.double(); qn = q.double().norm(dim=1)
    ref = torch.where((qn[:, None] * vn[None, :]) > 0, dot.double() / (qn[:, None] * vn[None, :]), torch.zeros((), device="cuda", dtype=torch.float64)).float()
    err = (cos - ref).abs().max().item()
    st = stats.cpu().numpy().view(np.uint32).copy()
    cos32 = eng.dense_scan(qd, eng._stats(B), "fp32")
    err32 = (cos - cos32).abs().max().item()
    dec = lambda e: np.array([((~e) & 0xFFFFFFFF) if not (e & 0x80000000) else (e & 0x7FFFFFFF)], np.uint32).view(np.float32)[0]
    mm_ok = all(dec(int(st[b, 0])) == cos[b].min().item() and dec(int(st[b, 1])) == cos[b].max().item() for b in range(B))
    print(f"n={n} d={d} B={B}: max|gemm - torch_bf16|={err:.3e}  max|gemm - fp32 path|={err32:.3e}  minmax_ok={mm_ok}")
    ok = ok and err < 2e-5 and err32 < 1e-2 and mm_ok
print("ALL OK" if ok else "MISMATCH")
# timing at scale
n, d = 4_000_000, 384
v = torch.randn(n, d, device="cuda")
shard = hs.DeviceIndex("cuda:0", n); shard.set_dense(v); del v
eng = SearchEngine(shard, max_batch=256)
for B in (32, 64, 128, 256):
    q = torch.randn(B, d).numpy()
    qd = eng.upload_vectors(q).clone(); stats = eng._stats(B)
    for _ in range(3): eng.dense_scan(qd, stats, "bf16")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): eng.dense_scan(qd, stats, "bf16")
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    passes = (B + 127) // 128
    flops = 2.0 * B * n * d
    byts = passes * n * d * 2 + B * n * 4
    print(f"B={B}: {ms:.3f} ms  {B/ms*1e3:.0f} q/s  {flops/ms/1e9:.1f} TFLOP/s ({flops/ms/1e9/1402.2:.3f} of sustained bf16 peak)  "
          f"{byts/ms/1e6:.0f} GB/s ({byts/ms/1e6/6547.2:.3f} of HBM peak)")
