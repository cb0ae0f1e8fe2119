"""A/B harness for GEMM-kernel builds: run with REST_B200_LIB=<variant .so>; prints one JSON line.
Shapes are chosen so that both 128x128 and 128x64 tilings fill whole waves of 148 SMs (M = 37*128, N = 16*128)."""
import json
import os
import sys

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


def main():
    ctx = Context(0)
    out = {"lib": os.path.basename(os.environ.get("REST_B200_LIB", "default"))}
    M, N = 37 * 128, 16 * 128
    for K in [64, 256, 600, 1024, 4096]:
        a = ctx.empty(K * M); b = ctx.empty(K * N); c = ctx.empty(M * N)
        ctx.fill_linear(a, K * M, 1, 0, 1.0); ctx.fill_linear(b, K * N, 2, 0, 1.0)
        ms = best_ms(lambda: ctx.dgemm("T", "N", M, N, K, 1.0, a, K, b, K, 0.0, c, M))
        out[f"TN_k{K}"] = round(2.0 * M * N * K / ms / 1e9, 2)
    Mb = 37 * 128 * 8
    for K in [600, 608, 1800]:
        a = ctx.empty(K * Mb); b = ctx.empty(K * N); c = ctx.empty(Mb * N)
        ctx.fill_linear(a, K * Mb, 1, 0, 1.0); ctx.fill_linear(b, K * N, 2, 0, 1.0)
        ms = best_ms(lambda: ctx.dgemm("T", "N", Mb, N, K, 1.0, a, K, b, K, 0.0, c, Mb))
        out[f"TNbig_k{K}"] = round(2.0 * Mb * N * K / ms / 1e9, 2)
        del a, b, c
    n = 8192
    a = ctx.empty(n * n); b = ctx.empty(n * n); c = ctx.empty(n * n)
    ctx.fill_linear(a, n * n, 1, 0, 1.0); ctx.fill_linear(b, n * n, 2, 0, 1.0)
    for ta, tb in [("T", "N"), ("N", "N")]:
        ms = best_ms(lambda: ctx.dgemm(ta, tb, n, n, n, 1.0, a, n, b, n, 0.0, c, n), reps=3, warm=1)
        out[f"{ta}{tb}_8192"] = round(2.0 * n ** 3 / ms / 1e9, 2)
    ms = best_ms(lambda: ctx.dsyrk("U", "N", n, n, 1.0, a, n, 0.0, c, n), reps=3, warm=1)
    out["syrk_8192"] = round(n * (n + 1.0) * n / ms / 1e9, 2)
    del a, b, c
    for name, nb, nx, no in [("A", 100, 400, 20), ("B", 264, 720, 21), ("C", 600, 1700, 60), ("D", 1800, 600, 180)]:
        sh = ShardedRI(ctx, nb, nx).fill_synthetic()
        cm = ctx.empty(nb * nb); ctx.fill_linear(cm, nb * nb, 3, 0, nb ** -0.5)
        mo = ctx.empty(nx * nb * nb)
        ct = cm[: nb * no].clone()
        k = ctx.empty(nb * nb)
        ms = best_ms(lambda: sh.ao2mo(cm, nb, cm, nb, out=mo), reps=3, warm=1)
        out[f"ao2mo_{name}"] = round(4.0 * nb ** 3 * nx / ms / 1e9, 2)
        ms = best_ms(lambda: sh.k(ct, no, out=k), reps=3, warm=1)
        out[f"k_{name}"] = round((2.0 * nb * nb * no + nb * (nb + 1.0) * no) * nx / ms / 1e9, 2)
        del sh, mo
        torch.cuda.empty_cache()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
