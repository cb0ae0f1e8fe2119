"""Find the smallest SYRK / GEMM shape whose repeated runs are not bitwise equal (TMA kernel, split-K)."""
import sys
import torch
sys.path.insert(0, ".")
from rest_tensors_b200.device import Context  # noqa: E402
from rest_tensors_b200._lib import lib  # noqa: E402

ctx = Context(0)
kmax = 108000
for n in (1800, 1544, 1032, 776):
    a = ctx.empty(n * kmax); ctx.fill_linear(a, n * kmax, 7, 0, 1.0)
    for k in (108000, 20000, 4000):
        s = lib.rb_gemm_plan_splits(n, n, k, 1, 1, ctx.num_sms)
        outs = []
        for rep in range(4):
            c = ctx.empty(n * n); c.fill_(float("nan"))
            ctx.dsyrk("U", "N", n, k, 1.0, a, n, 0.0, c, n)
            outs.append(torch.triu(c.view(n, n).t()).clone())    # upper triangle (row <= col) of the column-major matrix
        bad = [int((outs[r] != outs[0]).sum()) for r in range(1, 4)]
        nan = int(torch.isnan(outs[0]).sum())
        g = []
        for rep in range(2):
            c = ctx.empty(n * n)
            ctx.dgemm("N", "T", n, n, k, 1.0, a, n, a, n, 0.0, c, n)
            g.append(c.clone())
        print(f"n={n:5d} k={k:6d} splits={s:3d}  syrk mismatching elements vs run 0: {bad}  nan in upper: {nan}   full gemm mismatch: {int((g[0] != g[1]).sum())}", flush=True)
    del a
