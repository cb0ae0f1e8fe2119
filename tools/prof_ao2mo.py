"""Profiling target: config-C ao2mo (2 GEMM launches per call) on one GPU; run under ncu (see profiles/)."""
import sys
import torch
sys.path.insert(0, ".")
from rest_tensors_b200.device import Context, ShardedRI  # noqa: E402

nb, nx = (600, 1700) if len(sys.argv) < 3 else (int(sys.argv[1]), int(sys.argv[2]))
ctx = Context(0)
sh = ShardedRI(ctx, nb, nx).fill_synthetic()
c = ctx.empty(nb * nb); ctx.fill_linear(c, nb * nb, 3, 0, nb ** -0.5)
mo = ctx.empty(nx * nb * nb)
for _ in range(2):
    sh.ao2mo(c, nb, c, nb, out=mo)
torch.cuda.synchronize()
