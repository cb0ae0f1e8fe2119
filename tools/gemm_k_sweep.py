"""Per-tile overhead of the GEMM kernel: time vs K at fixed M=N (TN path). time = tiles/148 * (a*K + b)."""
import sys
sys.path.insert(0, ".")
import torch
from rest_tensors_b200.device import Context
ctx = Context(0)
M = N = 148 * 128 // 4 * 1  # 4736 -> 37x37 = 1369 tiles
M = N = 4736
rows = []
for K in [64, 128, 256, 512, 1024, 2048, 4096]:
    a = ctx.empty(K * M); b = ctx.empty(K * N); c = ctx.empty(M * N)
    ctx.fill_linear(a, K * M, 1, 0, 1.0); ctx.fill_linear(b, K * N, 2, 0, 1.0)
    for _ in range(2):
        ctx.dgemm("T", "N", M, N, K, 1.0, a, K, b, K, 0.0, c, M)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ctx.dgemm("T", "N", M, N, K, 1.0, a, K, b, K, 0.0, c, M); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    tiles = (M // 128) ** 2
    per_tile_us = best * 1e3 / (tiles / 148.0)
    rows.append((K, best, per_tile_us))
    print(f"K={K:5d}  {best:8.3f} ms  {2.0*M*N*K/best/1e9:7.2f} TFLOP/s  per-tile-slot {per_tile_us:7.2f} us", flush=True)
(k1, _, t1), (k2, _, t2) = rows[2], rows[-1]
a = (t2 - t1) / (k2 - k1); b = t1 - a * k1
print(f"fit: per-tile = {a*1e3:.3f} ns * K + {b:.2f} us overhead; ideal slope at 37.05 TF: {2*128*128/37.05e12*148*1e9:.3f} ns/K")
