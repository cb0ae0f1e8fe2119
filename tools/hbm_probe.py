"""HBM-bound kernels: achieved GB/s of algorithmic bytes on one GPU.  Usage: python tools/hbm_probe.py [out.json]"""
import json
import sys

import torch

sys.path.insert(0, ".")
from rest_tensors_b200.device import Context  # noqa: E402


def best_ms(fn, reps=7, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


def stream_ms(make_call, bytes_per_call, target=1 << 30):
    """per-call time with the calls issued back to back over distinct buffer sets (>= 1 GB in total, so nothing is served
    from L2) inside ONE event pair: the sustained rate of the kernel including its own launch gap, without the ~5 us an
    event pair around a single launch adds"""
    sets = int(max(2, min(64, -(-target // bytes_per_call))))
    calls = [make_call() for _ in range(sets)]
    best = None
    for it in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for fn in calls:
            fn()
        e1.record(); torch.cuda.synchronize()
        if it:
            t = e0.elapsed_time(e1) / sets
            best = t if best is None else min(best, t)
    return best


def main():
    ctx = Context(0)
    out = {"hbm_copy_probe_gbs": round(ctx.hbm_copy_probe(4 << 30, 10), 1)}
    a = torch.empty(1 << 28, dtype=torch.float64, device="cuda:0"); b = torch.empty_like(a)
    out["torch_copy_gbs"] = round(2 * a.numel() * 8 / best_ms(lambda: b.copy_(a)) / 1e6, 1)
    del a, b
    for n in (1800, 4000, 8000):
        np_ = n * (n + 1) // 2
        p = ctx.empty(np_); f = ctx.empty(n * n); g = ctx.empty(n * n); ctx.fill_linear(p, np_, 4, 0, 1.0)
        out[f"unpack_{n}_gbs"] = round((np_ + n * n) * 8 / best_ms(lambda: ctx.unpack_upper(p, n, f)) / 1e6, 1)
        out[f"pack_{n}_gbs"] = round(2 * np_ * 8 / best_ms(lambda: ctx.pack_upper(f, n, p)) / 1e6, 1)
        out[f"transpose_{n}_gbs"] = round(2 * n * n * 8 / best_ms(lambda: ctx.matrix_transpose(f, n, n, g)) / 1e6, 1)
        out[f"copy_mm_{n}_gbs"] = round(2 * n * n * 8 / best_ms(lambda: ctx.copy_mm(n, n, f, n, n, 0, 0, g, n, n, 0, 0)) / 1e6, 1)
        out[f"axpy_{n}_gbs"] = round(3 * n * n * 8 / best_ms(lambda: ctx.self_scaled_add(g, f, 0.5, n * n)) / 1e6, 1)
        del p, f, g

        def mk(kind):
            def make():
                pp = ctx.empty(np_); ff = ctx.empty(n * n); gg = ctx.empty(n * n)
                return {"unpack": lambda: ctx.unpack_upper(pp, n, ff), "pack": lambda: ctx.pack_upper(ff, n, pp),
                        "transpose": lambda: ctx.matrix_transpose(ff, n, n, gg),
                        "copy_mm": lambda: ctx.copy_mm(n, n, ff, n, n, 0, 0, gg, n, n, 0, 0)}[kind]
            return make
        for kind, nbytes in (("unpack", (np_ + n * n) * 8), ("pack", 2 * np_ * 8), ("transpose", 2 * n * n * 8), ("copy_mm", 2 * n * n * 8)):
            out[f"{kind}_{n}_stream_gbs"] = round(nbytes / stream_ms(mk(kind), nbytes) / 1e6, 1)
    # the flat 256-bit copy on the micro-benchmark's footprint (1 GiB in, 1 GiB out): separates kernel structure from size effects
    nn = 11584
    big = ctx.empty(nn * nn); big2 = ctx.empty(nn * nn)
    out["copy_flat_1GiB_gbs"] = round(2 * nn * nn * 8 / best_ms(lambda: ctx.copy_mm(nn, nn, big, nn, nn, 0, 0, big2, nn, nn, 0, 0)) / 1e6, 1)
    out["transpose_1GiB_gbs"] = round(2 * nn * nn * 8 / best_ms(lambda: ctx.matrix_transpose(big, nn, nn, big2)) / 1e6, 1)
    del big, big2
    I, J, K = 600, 600, 400
    t = ctx.empty(I * J * K); u = ctx.empty(I * J * K); ctx.fill_linear(t, I * J * K, 5, 0, 1.0)
    for which, name in enumerate(["jik", "jki", "kji", "ikj"]):
        out[f"ri_transpose_{name}_gbs"] = round(2 * I * J * K * 8 / best_ms(lambda: ctx.ri_transpose(t, I, J, K, which, u)) / 1e6, 1)
    np_ = I * (I + 1) // 2
    out["ri_pack_symm_gbs"] = round(2 * np_ * K * 8 / best_ms(lambda: ctx.ri_pack_symm(t, I, K, u)) / 1e6, 1)
    xl, yl, zl = 500, 520, 300
    out["copy_rr_box_gbs"] = round(2 * xl * yl * zl * 8 / best_ms(lambda: ctx.copy_rr(xl, yl, zl, t, I, J, K, 50, 40, 30, u, I, J, K, 20, 10, 60)) / 1e6, 1)
    print(json.dumps(out), flush=True)
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
