"""Profiling target: rb_dsyev at n = 1800 (the Jacobi round kernel, replayed from the sweep graph); run under ncu."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from rest_tensors_b200.device import Context  # noqa: E402

n = 1800
ctx = Context(0)
a = ctx.empty(n * n); ctx.fill_linear(a, n * n, 71, 0, 1.0)
w = ctx.empty(n); z = ctx.empty(n * n)
ctx.dsyev("V", "L", n, a, n, w, z, n)
torch.cuda.synchronize()
