"""Profiling target: a 'T','N' GEMM made of thin (1-block-wide) tiles only (N = 8) next to a full-tile one (N = 128), same M and K."""
import sys
import torch
sys.path.insert(0, ".")
from rest_tensors_b200.device import Context  # noqa: E402

M, K = 190080, 1800
ctx = Context(0)
a = ctx.empty(K * M); ctx.fill_linear(a, K * M, 1, 0, 1.0)
for N in (8, 128):
    b = ctx.empty(K * N); ctx.fill_linear(b, K * N, 2, 0, 1.0)
    c = ctx.empty(M * N)
    for _ in range(2):
        ctx.dgemm("T", "N", M, N, K, 1.0, a, K, b, K, 0.0, c, M)
torch.cuda.synchronize()
