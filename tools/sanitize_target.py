"""Small, fast exercise of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from rest_tensors_b200.device import Context, ShardedRI  # noqa: E402

ctx = Context(0)
dev = "cuda:0"
rng = np.random.default_rng(0)
for ta, tb, (m, n, k) in [("T", "N", (300, 200, 77)), ("N", "N", (130, 260, 40)), ("N", "T", (257, 130, 9)), ("T", "T", (64, 140, 100)),
                          ("N", "N", (200, 130, 6000))]:
    ra, ca = (m, k) if ta == "N" else (k, m)
    rb, cb = (k, n) if tb == "N" else (n, k)
    a = torch.from_numpy(rng.standard_normal(ra * ca)).to(dev); b = torch.from_numpy(rng.standard_normal(rb * cb)).to(dev)
    c = ctx.empty(m * n)
    ctx.dgemm(ta, tb, m, n, k, 1.0, a, ra, b, rb, 0.0, c, m)
    ref = (a.view(ca, ra).t() if ta == "N" else a.view(ca, ra)) @ (b.view(cb, rb).t() if tb == "N" else b.view(cb, rb))
    err = float((c.view(n, m).t() - ref).abs().max() / ref.abs().max())
    assert err < 1e-12, (ta, tb, m, n, k, err)
a = torch.from_numpy(rng.standard_normal(300 * 50)).to(dev); c = ctx.empty(300 * 300); c.zero_()
ctx.dsyrk("U", "N", 300, 50, 1.0, a, 300, 0.0, c, 300)
nb, nx, no = 40, 70, 6
sh = ShardedRI(ctx, nb, nx).fill_synthetic()
cm = ctx.empty(nb * nb); ctx.fill_linear(cm, nb * nb, 3, 0, nb ** -0.5)
mo = sh.ao2mo(cm, nb, cm, nb)
d = sh.dp(cm); j = sh.j(d); kk = sh.k(cm[: nb * no].clone(), no)
n = 150; npk = n * (n + 1) // 2
p = ctx.empty(npk); f = ctx.empty(n * n); g = ctx.empty(n * n); ctx.fill_linear(p, npk, 4, 0, 1.0)
ctx.unpack_upper(p, n, f); ctx.pack_upper(f, n, p); ctx.matrix_transpose(f, n, n, g)
n = 151; npk = n * (n + 1) // 2
p = ctx.empty(npk); f = ctx.empty(n * n); g = ctx.empty(n * n); ctx.fill_linear(p, npk, 4, 0, 1.0)
ctx.unpack_upper(p, n, f); ctx.pack_upper(f, n, p); ctx.matrix_transpose(f, n, n, g)
t = ctx.empty(30 * 22 * 14); u = ctx.empty(30 * 22 * 14); ctx.fill_linear(t, 30 * 22 * 14, 5, 0, 1.0)
for w in range(4):
    ctx.ri_transpose(t, 30, 22, 14, w, u)
ctx.copy_rr(10, 12, 5, t, 30, 22, 14, 3, 2, 1, u, 30, 22, 14, 0, 4, 6)
ctx.self_scaled_add(g, f, 0.5, n * n)
# consumers of ri3mo: in-place panels, gathered boxes, weights, symmetric + general, beta accumulation
mo3 = ctx.empty(50 * 5 * 8); ctx.fill_linear(mo3, 50 * 5 * 8, 6, 0, 1.0)
wv = ctx.empty(5 * 8); ctx.fill_linear(wv, 40, 7, 0, 1.0)
gg = ctx.empty(40 * 40); gg.zero_()
ctx.ri_iajb(50, mo3, 50, 5, 8, (0, 5, 0, 8), mo3, 50, 5, 8, (0, 5, 0, 8), 0.0, gg, 40)
ctx.ri_iajb(50, mo3, 50, 5, 8, (1, 3, 2, 5), mo3, 50, 5, 8, (0, 5, 1, 6), 1.0, gg, 40)
ctx.ri_iajb(49, mo3[1:], 50, 5, 8, (1, 3, 2, 5), mo3[1:], 50, 5, 8, (1, 3, 2, 5), 0.0, gg, 40)
pp = ctx.empty(50 * 50)
ctx.ri_mo_pq(mo3, 50, 50, mo3, 50, 50, 5, 8, (0, 5, 0, 8), None, 0.0, pp, 50)
ctx.ri_mo_pq(mo3, 50, 50, mo3, 50, 50, 5, 8, (1, 3, 2, 5), wv, 0.0, pp, 50)
ctx.ri_mo_pq(mo3, 50, 20, mo3[20:], 50, 30, 5, 8, (0, 5, 0, 8), wv, 1.0, pp, 50)
# round 2: 32-byte layout kernels (multiples of 4, aligned), the bulk-tensor (TMA) path, stream-K products, ERIFold4 scatter
n = 256; npk = n * (n + 1) // 2
p = ctx.empty(npk); f = ctx.empty(n * n); g = ctx.empty(n * n); ctx.fill_linear(p, npk, 4, 0, 1.0)
ctx.unpack_upper(p, n, f); ctx.pack_upper(f, n, p); ctx.matrix_transpose(f, n, n, g); ctx.copy_mm(n, n, f, n, n, 0, 0, g, n, n, 0, 0)
assert torch.equal(f.view(n, n), f.view(n, n).t())
t = ctx.empty(64 * 72 * 20); u = ctx.empty(64 * 72 * 20); ctx.fill_linear(t, t.numel(), 5, 0, 1.0)
for path in (0, 1):
    ctx.set_layout_path(path)
    for w in range(4):
        ctx.ri_transpose(t, 64, 72, 20, w, u)
    ctx.copy_rr(32, 40, 10, t, 64, 72, 20, 4, 8, 2, u, 64, 72, 20, 8, 4, 6)
ctx.set_layout_path(0)
ctx.self_scaled_add(g, f, 0.5, n * n)
for (m, nn, k, tri) in [(500, 500, 500, 0), (264, 264, 4000, 1), (129, 300, 777, 0)]:
    a = ctx.empty(max(m, nn) * k); b = ctx.empty(max(m, nn) * k); c = ctx.empty(m * nn)
    ctx.fill_linear(a, a.numel(), 8, 0, 1.0); ctx.fill_linear(b, b.numel(), 9, 0, 1.0); c.zero_()
    if tri:
        ctx.dsyrk("U", "N", m, k, 1.0, a, m, 0.0, c, m)
    else:
        ctx.dgemm("N", "N", m, nn, k, 1.0, a, m, b, k, 0.0, c, m)
        ref = a[: m * k].view(k, m).t() @ b[: k * nn].view(nn, k).t()
        assert float((c.view(nn, m).t() - ref).abs().max() / ref.abs().max()) < 1e-12
dim = 12; npair = dim * (dim + 1) // 2
eri = ctx.empty(npair * npair); eri.zero_()
blk = ctx.empty(5 * 7 * 4 * 6); ctx.fill_linear(blk, blk.numel(), 10, 0, 1.0)
ctx.erifold4_chunk_copy(eri, npair, npair, npair, ((0, 5), (5, 12), (2, 6), (6, 12)), blk, 1)
ctx.erifold4_chunk_copy(eri, npair, npair, npair, ((0, 5), (0, 7), (0, 4), (0, 6)), blk, 0)
# d_P + J from one read (persistent cooperative kernel: LDGSTS ring, named-barrier hand-off to the exchange warps, sentinel-valued exchange)
import os  # noqa: E402
os.environ["REST_B200_DPJ_FUSED"] = "1"
for nb2, nx2 in [(40, 70), (128, 33), (264, 40)]:
    sh2 = ShardedRI(ctx, nb2, nx2).fill_synthetic()
    dm2 = ctx.empty(nb2 * nb2); ctx.fill_linear(dm2, nb2 * nb2, 4, 0, 1.0 / nb2)
    d_ref = sh2.dp(dm2); j_ref = sh2.j(d_ref, reduce=False)
    d1, j1 = sh2.dp_j(dm2, reduce=False)
    assert float((d1 - d_ref).abs().max()) <= 1e-13 * float(d_ref.abs().max())
    assert float((j1 - j_ref).abs().max()) <= 1e-13 * float(j_ref.abs().max())
del os.environ["REST_B200_DPJ_FUSED"]
torch.cuda.synchronize()
print("sanitize target ok")
