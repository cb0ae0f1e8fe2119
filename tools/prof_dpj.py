"""d_P + J: two passes (rb_ri_dp, rb_ri_j) against the fused single pass (rb_ri_dp_j); CUDA events, back to back, per config."""
import sys
import torch
sys.path.insert(0, ".")
import os
os.environ["REST_B200_DPJ_FUSED"] = "1"
from rest_tensors_b200.device import Context, ShardedRI  # noqa: E402
ctx = Context(0)
for nb, nx in [(600, 1700), (700, 800), (800, 600), (900, 400), (1000, 400), (264, 720)]:
    sh = ShardedRI(ctx, nb, nx).fill_synthetic()
    dm = ctx.empty(nb * nb); ctx.fill_linear(dm, nb * nb, 4, 0, 1.0 / nb)
    d = ctx.empty(nx); j = ctx.empty(nb * nb)

    def two():
        sh.dp(dm, out=d); sh.j(d, out=j, reduce=False)

    def one():
        sh.dp_j(dm, out_d=d, out_j=j, reduce=False)
    res = []
    for fn in (two, one):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / reps)
    gb = nb * nb * nx * 8 / 1e9
    print(f"nb={nb} nx={nx} ({gb:.2f} GB): two passes {res[0]*1e3:.1f} us ({2*gb/res[0]:.0f} GB/s of 2 reads)  fused {res[1]*1e3:.1f} us "
          f"({gb/res[1]:.0f} GB/s of 1 read)  speed-up {res[0]/res[1]:.2f}x", flush=True)
