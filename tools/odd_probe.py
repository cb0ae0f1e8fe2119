"""Odd leading dimensions cannot be described by TMA (rows must start 16-byte aligned): how slow is that path?"""
import sys
import torch
sys.path.insert(0, ".")
from rest_tensors_b200.device import Context, ShardedRI  # noqa: E402


def best_ms(fn, reps=3, warm=1):
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
for n in (2048, 2049):
    a = ctx.empty(n * n); b = ctx.empty(n * n); c = ctx.empty(n * n)
    ctx.fill_linear(a, n * n, 1, 0, 1.0); ctx.fill_linear(b, n * n, 2, 0, 1.0)
    for ta, tb in (("T", "N"), ("N", "N")):
        ms = best_ms(lambda: ctx.dgemm(ta, tb, n, n, n, 1.0, a, n, b, n, 0.0, c, n))
        print(f"dgemm {ta}{tb} n={n}: {2.0 * n ** 3 / ms / 1e9:.2f} TFLOP/s", flush=True)
for nb, nx, no in ((600, 400, 60), (601, 400, 60)):
    sh = ShardedRI(ctx, nb, nx).fill_synthetic()
    c = ctx.empty(nb * nb); ctx.fill_linear(c, nb * nb, 3, 0, nb ** -0.5)
    mo = ctx.empty(nx * nb * nb); k = ctx.empty(nb * nb); ct = c[: nb * no].clone()
    ms = best_ms(lambda: sh.ao2mo(c, nb, c, nb, out=mo))
    print(f"ao2mo nb={nb} nx={nx}: {4.0 * nb ** 3 * nx / ms / 1e9:.2f} TFLOP/s", flush=True)
    ms = best_ms(lambda: sh.k(ct, no, out=k))
    print(f"K     nb={nb} nx={nx}: {(2.0 * nb * nb * no + nb * (nb + 1.0) * no) * nx / ms / 1e9:.2f} TFLOP/s", flush=True)
for nb, nx in ((600, 400), (601, 400)):
    sh = ShardedRI(ctx, nb, nx).fill_synthetic()
    dm = ctx.empty(nb * nb); ctx.fill_linear(dm, nb * nb, 5, 0, 1.0)
    d = ctx.empty(nx); j = ctx.empty(nb * nb)
    ms = best_ms(lambda: sh.dp(dm, out=d), reps=5)
    print(f"d_P nb={nb}: {nb * nb * nx * 8 / ms / 1e6:.0f} GB/s", flush=True)
    ms = best_ms(lambda: sh.j(d, out=j), reps=5)
    print(f"J   nb={nb}: {nb * nb * nx * 8 / ms / 1e6:.0f} GB/s", flush=True)
