"""How expensive are thin edge tiles?  GEMM time vs N around tile boundaries, with the thin-edge-tile loads on (gemm path 0)
and off (path 2: edge tiles stream full zero-filled TMA boxes).  K-major ('T','N': the ao2mo layout) and MN-major B ('T','T')."""
import json, sys
import torch
sys.path.insert(0, ".")
from rest_tensors_b200.device import Context  # noqa: E402


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
CASES = [("N", 37888 * 2, 1800, [1792, 1800, 1808, 1856]), ("N", 190080, 264, [256, 264, 272, 288, 320]),
         ("N", 37888 * 8, 600, [512, 520, 600]), ("T", 190080, 264, [256, 264, 272, 320]),
         ("N", 1700 * 600, 600, [60])]  # last: the shape class of K's first GEMM (N = nocc)
for tb, M, K, Ns in CASES:
    a = ctx.empty(K * M); ctx.fill_linear(a, K * M, 1, 0, 1.0)
    for N in Ns:
        b = ctx.empty(K * N); c = ctx.empty(M * N); ctx.fill_linear(b, K * N, 2, 0, 1.0)
        ldb = K if tb == "N" else N
        row = {}
        for path, name in [(0, "thin"), (2, "full")]:
            ctx.set_gemm_path(path)
            ms = best_ms(lambda: ctx.dgemm("T", tb, M, N, K, 1.0, a, K, b, ldb, 0.0, c, M))
            row[name] = {"ms": round(ms, 3), "tflops": round(2.0 * M * N * K / ms / 1e9, 2)}
            if path == 0:
                ref = c.clone()
            else:
                row["bitwise_equal"] = bool(torch.equal(ref, c))
        ctx.set_gemm_path(0)
        out[f"T{tb}_M{M}_K{K}_N{N}"] = row
        del b, c, ref
    del a
    torch.cuda.empty_cache()
for k, v in out.items():
    print(k, json.dumps(v))
json.dump(out, open("gpurun_out/thin_tile_probe.json", "w"), indent=1)
