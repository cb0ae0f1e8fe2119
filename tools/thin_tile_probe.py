"""How expensive are thin edge tiles?  TN GEMM time vs N around tile boundaries."""
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
for M, K, Ns in [(37888 * 2, 1800, [1792, 1800, 1808, 1856, 1920]), (190080, 264, [256, 264, 272, 288, 320, 384]), (37888 * 8, 600, [512, 520, 600, 608, 640])]:
    a = ctx.empty(K * M); ctx.fill_linear(a, K * M, 1, 0, 1.0)
    for N in Ns:
        b = ctx.empty(K * N); c = ctx.empty(M * N); ctx.fill_linear(b, K * N, 2, 0, 1.0)
        ms = best_ms(lambda: ctx.dgemm("T", "N", M, N, K, 1.0, a, K, b, K, 0.0, c, M))
        out[f"M{M}_K{K}_N{N}"] = {"ms": round(ms, 3), "tflops": round(2.0 * M * N * K / ms / 1e9, 2), "ms_per_128cols": round(ms / (N / 128.0), 4)}
        del b, c
    del a
    torch.cuda.empty_cache()
for k, v in out.items():
    print(k, v)
