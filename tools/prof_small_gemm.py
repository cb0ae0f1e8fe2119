"""Profiling target: small square dgemm / dsyrk (config E lower half) for an ncu launch list."""
import sys
import torch
sys.path.insert(0, ".")
from rest_tensors_b200.device import Context  # noqa: E402
ctx = Context(0)
for n in (500, 1000, 2000):
    a = ctx.empty(n * n); b = ctx.empty(n * n); c = ctx.empty(n * n)
    ctx.fill_linear(a, n * n, 5, 0, 1.0); ctx.fill_linear(b, n * n, 6, 0, 1.0)
    for _ in range(3):
        ctx.dgemm("N", "N", n, n, n, 1.0, a, n, b, n, 0.0, c, n)
        ctx.dsyrk("U", "N", n, n, 1.0, a, n, 0.0, c, n)
torch.cuda.synchronize()
