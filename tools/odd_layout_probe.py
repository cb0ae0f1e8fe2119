import sys, torch
sys.path.insert(0, ".")
from rest_tensors_b200.device import Context
def best_ms(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
ctx = Context(0)
for n in (8000, 8001):
    np_ = n * (n + 1) // 2
    p = ctx.empty(np_); f = ctx.empty(n * n); g = ctx.empty(n * n); ctx.fill_linear(p, np_, 4, 0, 1.0)
    print(n, "unpack", round((np_ + n * n) * 8 / best_ms(lambda: ctx.unpack_upper(p, n, f)) / 1e6),
          "pack", round(2 * np_ * 8 / best_ms(lambda: ctx.pack_upper(f, n, p)) / 1e6),
          "transpose", round(2 * n * n * 8 / best_ms(lambda: ctx.matrix_transpose(f, n, n, g)) / 1e6),
          "copy_mm", round(2 * n * n * 8 / best_ms(lambda: ctx.copy_mm(n, n, f, n, n, 0, 0, g, n, n, 0, 0)) / 1e6), flush=True)
for I, J, K in ((600, 600, 400), (601, 601, 401)):
    t = ctx.empty(I * J * K); u = ctx.empty(I * J * K); ctx.fill_linear(t, I * J * K, 5, 0, 1.0)
    print(I, [round(2 * I * J * K * 8 / best_ms(lambda: ctx.ri_transpose(t, I, J, K, w, u)) / 1e6) for w in range(4)],
          "pack_symm", round(2 * (I * (I + 1) // 2) * K * 8 / best_ms(lambda: ctx.ri_pack_symm(t, I, K, u)) / 1e6), flush=True)
