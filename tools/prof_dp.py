"""Profiling target: config-C d_P (rb_gemv_t_vec_kernel) and J (rb_gemv_n_kernel), one pass over ri3ao each; run under ncu."""
import sys
import torch
sys.path.insert(0, ".")
from rest_tensors_b200.device import Context, ShardedRI  # noqa: E402

nb, nx = 600, 1700
ctx = Context(0)
sh = ShardedRI(ctx, nb, nx).fill_synthetic()
dm = ctx.empty(nb * nb); ctx.fill_linear(dm, nb * nb, 3, 0, 1.0)
for _ in range(2):
    d = sh.dp(dm)
    j = sh.j(d, reduce=False)
torch.cuda.synchronize()
