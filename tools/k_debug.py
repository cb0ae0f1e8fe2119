"""Diagnose the K additivity failure at the config-D shard shape: determinism of repeated builds, where and by how much the
half-shard sum differs from the whole."""
import sys
import torch
sys.path.insert(0, ".")
from rest_tensors_b200.device import Context, ShardedRI  # noqa: E402

nb, nx, no = 1800, 600, 180
ctx = Context(0)
if len(sys.argv) > 1:
    ctx.set_gemm_path(int(sys.argv[1]))   # 1 = the generic plain-load kernel
ri = ShardedRI(ctx, nb, nx).fill_synthetic()
n2 = nb * nb
c = ctx.empty(n2); ctx.fill_linear(c, n2, 3, 0, nb ** -0.5)
cm = c.view(nb, nb).t()
ct = (cm[:, :no] * (2.0 ** 0.5)).t().contiguous().reshape(-1)
k1 = ri.k(ct, no, reduce=False).clone()
k2 = ri.k(ct, no, reduce=False).clone()
print("whole twice bitwise equal:", bool(torch.equal(k1, k2)), "max |diff|", float((k1 - k2).abs().max()))
half = (nx // 2) // 8 * 8 + 4
lo = ShardedRI(ctx, nb, half, data=ri.data[: n2 * half])
hi = ShardedRI(ctx, nb, nx - half, data=ri.data[n2 * half:])
for rep in range(3):
    kl = lo.k(ct, no, reduce=False).clone(); kh = hi.k(ct, no, reduce=False).clone()
    diff = (kl + kh - k1).abs().view(nb, nb)          # [col, row]
    big = (diff > 1e-9 * float(k1.abs().max())).nonzero()
    print(f"rep {rep}: max diff {float(diff.max()):.3e} (max |K| {float(k1.abs().max()):.3e}); elements beyond 1e-9: {big.shape[0]}")
    if big.shape[0]:
        cols = sorted(set((big[:, 0] // 16).tolist())); rows = sorted(set((big[:, 1] // 16).tolist()))
        print("   16-col blocks:", cols[:20], " 16-row blocks:", rows[:20])
        print("   first:", big[0].tolist(), "lo+hi", float((kl + kh).view(nb, nb)[big[0][0], big[0][1]]), "whole", float(k1.view(nb, nb)[big[0][0], big[0][1]]))
    # which side is wrong?  compare each half with a two-chunk build of itself through the oracle-free identity K(lo) = K(lo_a) + K(lo_b)
    q = half // 2
    la = ShardedRI(ctx, nb, q, data=ri.data[: n2 * q]); lb = ShardedRI(ctx, nb, half - q, data=ri.data[n2 * q: n2 * half])
    d2 = (la.k(ct, no, reduce=False) + lb.k(ct, no, reduce=False) - kl).abs()
    print(f"   K(lo) vs its own halves: max diff {float(d2.max()):.3e}")
