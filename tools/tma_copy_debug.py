"""Where does the bulk-tensor copy differ from the expected sub-box copy?  (odd x extent)"""
import sys
import torch
sys.path.insert(0, ".")
from rest_tensors_b200.device import Context  # noqa: E402

ctx = Context(0)
fx, fy, fz, tx, ty, tz = 600, 520, 12, 640, 530, 14
f = ctx.empty(fx * fy * fz); ctx.fill_linear(f, f.numel(), 23, 0, 1.0)
for (xl, yl, zl, fs, ts) in [(501, 500, 10, (50, 10, 1), (20, 4, 2)), (501, 16, 1, (50, 10, 1), (20, 4, 2)), (250, 16, 1, (50, 10, 1), (20, 4, 2)),
                             (251, 300, 1, (0, 0, 0), (0, 0, 0)), (499, 300, 1, (0, 0, 0), (0, 0, 0)), (254, 300, 1, (0, 0, 0), (0, 0, 0))]:
    t = ctx.empty(tx * ty * tz); ctx.fill_linear(t, t.numel(), 24, 0, 1.0)
    ref = t.clone()
    ref.view(tz, ty, tx)[ts[2]:ts[2] + zl, ts[1]:ts[1] + yl, ts[0]:ts[0] + xl] = f.view(fz, fy, fx)[fs[2]:fs[2] + zl, fs[1]:fs[1] + yl, fs[0]:fs[0] + xl]
    before = ctx.tma_layout_launches
    ctx.copy_rr(xl, yl, zl, f, fx, fy, fz, fs[0], fs[1], fs[2], t, tx, ty, tz, ts[0], ts[1], ts[2])
    bad = (t != ref).view(tz, ty, tx).nonzero()
    print((xl, yl, zl), "tma" if ctx.tma_layout_launches > before else "plain", "mismatches", bad.shape[0])
    if bad.shape[0]:
        xs = sorted(set((bad[:, 2] - ts[0]).tolist()))
        ys = sorted(set((bad[:, 1] - ts[1]).tolist()))
        print("   x (box coords):", xs[:12], "... y:", ys[:12], "first", bad[0].tolist(), "got", float(t.view(tz, ty, tx)[tuple(bad[0].tolist())]),
              "want", float(ref.view(tz, ty, tx)[tuple(bad[0].tolist())]))
