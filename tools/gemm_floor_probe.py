"""Fixed cost of one rb_dgemm call: tiny products issued back to back (per-call time), and under ncu the kernel duration itself."""
import sys
import torch
sys.path.insert(0, ".")
from rest_tensors_b200.device import Context  # noqa: E402
ctx = Context(0)
a = ctx.empty(4096 * 4096); b = ctx.empty(4096 * 4096); c = ctx.empty(4096 * 4096)
ctx.fill_linear(a, a.numel(), 5, 0, 1.0); ctx.fill_linear(b, b.numel(), 6, 0, 1.0)
for (m, n, k) in [(128, 128, 32), (128, 128, 320), (128, 128, 3200), (1024, 1024, 32), (1536, 1536, 32), (1536, 1536, 320), (500, 500, 500), (1000, 1000, 1000)]:
    for _ in range(5):
        ctx.dgemm("N", "N", m, n, k, 1.0, a, m, b, k, 0.0, c, m)
    torch.cuda.synchronize()
    reps = 200
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ctx.dgemm("N", "N", m, n, k, 1.0, a, m, b, k, 0.0, c, m)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    ideal = 2.0 * m * n * k / 37.1e12 * 1e6
    print(f"dgemm {m}x{n}x{k}: {us:.1f} us per call back to back (DMMA time at peak {ideal:.2f} us)", flush=True)
