"""Profiling target for the layout kernels at n (default 8000): unpack, pack, transpose, 3 calls each (run under ncu --set full -k regex)."""
import sys
import torch
sys.path.insert(0, ".")
from rest_tensors_b200.device import Context  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
ctx = Context(0)
npk = n * (n + 1) // 2
p = ctx.empty(npk); f = ctx.empty(n * n); g = ctx.empty(n * n)
ctx.fill_linear(p, npk, 4, 0, 1.0)
for _ in range(3):
    ctx.unpack_upper(p, n, f)
    ctx.pack_upper(f, n, p)
    ctx.matrix_transpose(f, n, n, g)
torch.cuda.synchronize()
print("ok")
