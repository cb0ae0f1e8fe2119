"""Quick on-box measurements (not the bench): FP64 pipe peaks, HBM copy, GEMM sweep, RI ops at configs A/B/C.
Usage: python tools/probe.py [out.json]"""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from rest_tensors_b200.device import Context, ShardedRI  # noqa: E402


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    out = {}
    ctx = Context(0)
    out["num_sms"] = ctx.num_sms
    for kind, name in [(0, "dmma"), (1, "dfma")]:
        vals = [ctx.fp64_peak_probe(kind, 8192)[0] for _ in range(3)]
        out[f"fp64_peak_{name}_tflops"] = max(vals)
    out["hbm_copy_gbs"] = ctx.hbm_copy_probe(4 << 30, 10)
    print(json.dumps(out), flush=True)
    # GEMM sweep
    gem = {}
    for n in [512, 1024, 2048, 4096, 8192]:
        a = ctx.empty(n * n); b = ctx.empty(n * n); c = ctx.empty(n * n)
        ctx.fill_linear(a, n * n, 1, 0, 1.0); ctx.fill_linear(b, n * n, 2, 0, 1.0)
        for ta, tb in [("T", "N"), ("N", "N"), ("N", "T")]:
            med, best = timed(lambda: ctx.dgemm(ta, tb, n, n, n, 1.0, a, n, b, n, 0.0, c, n), reps=3, warm=1)
            gem[f"dgemm_{ta}{tb}_{n}"] = {"ms": med, "tflops": 2.0 * n ** 3 / (best * 1e-3) / 1e12}
        med, best = timed(lambda: ctx.dsyrk("U", "N", n, n, 1.0, a, n, 0.0, c, n), reps=3, warm=1)
        gem[f"dsyrk_UN_{n}"] = {"ms": med, "tflops": n * (n + 1.0) * n / (best * 1e-3) / 1e12}
        print(n, {k: round(v["tflops"], 2) for k, v in gem.items() if k.endswith(f"_{n}")}, flush=True)
        del a, b, c
    out["gemm"] = gem
    # RI ops
    ri_out = {}
    cfgs = [("A", 100, 400, 20), ("B", 264, 720, 21), ("C", 600, 1700, 60)]
    if "--with-d" in sys.argv:
        cfgs.append(("D", 1800, 600, 180))
    for name, nb, nx, no in cfgs:
        sh = ShardedRI(ctx, nb, nx).fill_synthetic()
        c = ctx.empty(nb * nb); ctx.fill_linear(c, nb * nb, 3, 0, nb ** -0.5)
        mo = ctx.empty(nx * nb * nb)
        dm = ctx.empty(nb * nb); ctx.fill_linear(dm, nb * nb, 5, 0, 1.0)
        ct = c[: nb * no].clone()
        d = ctx.empty(nx); j = ctx.empty(nb * nb); k = ctx.empty(nb * nb)
        r = {}
        med, best = timed(lambda: sh.ao2mo(c, nb, c, nb, out=mo), reps=3, warm=1)
        r["ao2mo_ms"] = med; r["ao2mo_tflops"] = 4.0 * nb ** 3 * nx / (best * 1e-3) / 1e12
        med, best = timed(lambda: sh.dp(dm, out=d)); r["dp_ms"] = med; r["dp_gbs"] = nb * nb * nx * 8 / (best * 1e-3) / 1e9
        med, best = timed(lambda: sh.j(d, out=j)); r["j_ms"] = med; r["j_gbs"] = nb * nb * nx * 8 / (best * 1e-3) / 1e9
        med, best = timed(lambda: sh.k(ct, no, out=k), reps=3, warm=1)
        r["k_ms"] = med; r["k_tflops"] = (2.0 * nb * nb * no + nb * (nb + 1.0) * no) * nx / (best * 1e-3) / 1e12
        ri_out[name] = r
        print(name, {kk: round(v, 3) for kk, v in r.items()}, flush=True)
        del sh, mo
        torch.cuda.empty_cache()
    out["ri"] = ri_out
    # layout kernels at n = 8000
    n = 8000; np_ = n * (n + 1) // 2
    p = ctx.empty(np_); f = ctx.empty(n * n); ctx.fill_linear(p, np_, 4, 0, 1.0)
    med, best = timed(lambda: ctx.unpack_upper(p, n, f)); out["unpack_8000_gbs"] = (np_ + n * n) * 8 / (best * 1e-3) / 1e9
    med, best = timed(lambda: ctx.pack_upper(f, n, p)); out["pack_8000_gbs"] = 2 * np_ * 8 / (best * 1e-3) / 1e9
    g = ctx.empty(n * n)
    med, best = timed(lambda: ctx.matrix_transpose(f, n, n, g)); out["transpose_8000_gbs"] = 2 * n * n * 8 / (best * 1e-3) / 1e9
    med, best = timed(lambda: ctx.copy_mm(n, n, f, n, n, 0, 0, g, n, n, 0, 0)); out["copy_8000_gbs"] = 2 * n * n * 8 / (best * 1e-3) / 1e9
    med, best = timed(lambda: ctx.self_scaled_add(g, f, 0.5, n * n)); out["axpy_8000_gbs"] = 3 * n * n * 8 / (best * 1e-3) / 1e9
    print({k: v for k, v in out.items() if k.endswith("gbs")}, flush=True)
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    path = args[0] if args else "gpurun_out/probe.json"
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
