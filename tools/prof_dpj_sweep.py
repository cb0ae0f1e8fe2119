"""Fused d_P + J at config C for several ring shapes (REST_B200_DPJ_S slabs per block, REST_B200_DPJ_LAG blocks between the stages)."""
import os, sys
import torch
sys.path.insert(0, ".")
os.environ["REST_B200_DPJ_FUSED"] = "1"
from rest_tensors_b200.device import Context, ShardedRI  # noqa: E402
ctx = Context(0)
nb, nx = 600, 1700
sh = ShardedRI(ctx, nb, nx).fill_synthetic()
dm = ctx.empty(nb * nb); ctx.fill_linear(dm, nb * nb, 4, 0, 1.0 / nb)
d = ctx.empty(nx); j = ctx.empty(nb * nb)
for S, LAG, dbg in [(2, 3, 0), (2, 2, 0), (2, 3, 1), (1, 3, 0), (1, 4, 0), (1, 5, 0), (1, 6, 0), (1, 7, 0)]:
    os.environ["REST_B200_DPJ_S"] = str(S); os.environ["REST_B200_DPJ_LAG"] = str(LAG); os.environ["REST_B200_DPJ_DEBUG"] = str(dbg)
    for _ in range(2):
        sh.dp_j(dm, out_d=d, out_j=j, reduce=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        sh.dp_j(dm, out_d=d, out_j=j, reduce=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    R = min(10, 204800 // (S * 2436 * 8))
    print(f"debug={dbg} S={S} R={R} LAG={LAG}: {ms*1e3:.0f} us  {4.896/ms:.2f} TB/s", flush=True)
