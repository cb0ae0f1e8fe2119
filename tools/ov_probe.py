"""occ-vir (north-star form) ao2mo rates at configs B, C, D-shard: flop = (2 no nb^2 + 2 no nb nv) nx."""
import json, sys
import torch
sys.path.insert(0, ".")
from rest_tensors_b200.device import Context, ShardedRI  # noqa: E402


def best_ms(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


ctx = Context(0)
out = {}
for name, nb, nx, no in [("B", 264, 720, 21), ("C", 600, 1700, 60), ("D", 1800, 600, 180)]:
    nv = nb - no
    sh = ShardedRI(ctx, nb, nx).fill_synthetic()
    c = ctx.empty(nb * nb); ctx.fill_linear(c, nb * nb, 3, 0, nb ** -0.5)
    cl, cr = c[: nb * no], c[nb * no:]
    mo = ctx.empty(nx * no * nv)
    ms = best_ms(lambda: sh.ao2mo(cl, no, cr, nv, out=mo))
    flop = (2.0 * no * nb * nb + 2.0 * no * nb * nv) * nx
    out[f"ov_{name}"] = {"ms": round(ms, 3), "tflops": round(flop / ms / 1e9, 2), "ri3ao_gbs": round(nb * nb * nx * 8 / ms / 1e6, 1)}
    # the other association: (A_P C_vir) first costs 2 nv nb^2 + 2 no nb nv -- for reference only
    del sh, mo
    torch.cuda.empty_cache()
print(json.dumps(out))
