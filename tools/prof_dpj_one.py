import os, sys
import torch
sys.path.insert(0, ".")
from rest_tensors_b200.device import Context, ShardedRI  # noqa: E402
os.environ["REST_B200_DPJ_FUSED"] = "1"
ctx = Context(0)
nb, nx = 600, 1700
sh = ShardedRI(ctx, nb, nx).fill_synthetic()
dm = ctx.empty(nb * nb); ctx.fill_linear(dm, nb * nb, 4, 0, 1.0 / nb)
d = ctx.empty(nx); j = ctx.empty(nb * nb)
for _ in range(3):
    sh.dp_j(dm, out_d=d, out_j=j, reduce=False)
torch.cuda.synchronize()
print("ok")
