"""rb_ri_dp_j: d_P and J from ONE pass over ri3ao (persistent cooperative kernel, rest_tensors_b200/csrc/rb_dpj.cu) against the two-pass
path (rb_ri_dp + rb_ri_j = the reference's _dgemv 'T' / 'N' composition, /root/reference/src/ri.rs + matrix_blas_lapack.rs) and the oracle."""
from __future__ import annotations

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _fused_kernel_on(monkeypatch):
    """the single-pass kernel is opt-in (measured slower than the two passes); these tests exercise it"""
    monkeypatch.setenv("REST_B200_DPJ_FUSED", "1")


def _setup(ctx, nb, nx, seed=7):
    from rest_tensors_b200.device import ShardedRI
    sh = ShardedRI(ctx, nb, nx).fill_synthetic()
    dm = ctx.empty(nb * nb); ctx.fill_linear(dm, nb * nb, seed, 0, 1.0 / nb)
    return sh, dm


@pytest.mark.parametrize("nb,nx", [(100, 400), (264, 720), (600, 300), (40, 70), (10, 3), (2, 1), (64, 1), (128, 33), (800, 40), (900, 24),
                                   (45, 77), (1000, 12), (1200, 6), (36, 2000)])
def test_fused_dp_j_matches_two_passes(ctx, nb, nx):
    """same d_P and J as the two GEMV passes to summation-order accuracy (1e-13 of the largest element), bit-identical across
    repeated launches and with every workspace poisoned in between; odd nb and nb > 1100 take the fallback and are then identical."""
    sh, dm = _setup(ctx, nb, nx)
    d_ref = sh.dp(dm); j_ref = sh.j(d_ref, reduce=False)
    d1, j1 = sh.dp_j(dm, reduce=False)
    ctx.poison_workspaces()
    d2, j2 = sh.dp_j(dm, reduce=False)
    assert torch.equal(d1, d2) and torch.equal(j1, j2)
    assert bool(torch.isfinite(d1).all()) and bool(torch.isfinite(j1).all())
    sd, sj = float(d_ref.abs().max()), float(j_ref.abs().max())
    assert float((d1 - d_ref).abs().max()) <= 1e-13 * sd, float((d1 - d_ref).abs().max()) / sd
    assert float((j1 - j_ref).abs().max()) <= 1e-13 * sj, float((j1 - j_ref).abs().max()) / sj
    if nb % 2 == 1 or nb > 1100:
        assert torch.equal(d1, d_ref) and torch.equal(j1, j_ref)


def test_fused_dp_j_vs_oracle(ctx, oracle_blas):
    nb, nx = 72, 130
    ri = oracle_blas.fill_linear(nb * nb * nx, 2)              # not symmetric: nothing in the kernel may assume it
    dm = oracle_blas.fill_linear(nb * nb, 5, scale=1.0 / nb)
    dev = lambda a: torch.from_numpy(a).to("cuda:0")  # noqa: E731
    d = ctx.empty(nx); j = ctx.empty(nb * nb)
    ctx.ri_dp_j(dev(ri), dev(dm), d, j, nb, nx)
    d_ref = oracle_blas.ri_dp(ri, dm, nb, nx)
    j_ref = oracle_blas.ri_j(ri, d_ref, nb, nx)
    assert np.max(np.abs(d.cpu().numpy() - d_ref)) <= 1e-10 * np.max(np.abs(d_ref))
    assert np.max(np.abs(j.cpu().numpy() - j_ref)) <= 1e-10 * np.max(np.abs(j_ref))


def test_fused_dp_j_linearity_at_full_size(ctx):
    """config C shape on a slab subset: d_P and J are linear in D (size-independent property), and the fused pass agrees with the
    two-pass path at nb = 600"""
    nb, nx = 600, 512
    sh, dm = _setup(ctx, nb, nx)
    dm2 = ctx.empty(nb * nb); ctx.fill_linear(dm2, nb * nb, 9, 0, 1.0 / nb)
    d_a, j_a = sh.dp_j(dm, reduce=False)
    d_b, j_b = sh.dp_j(dm2, reduce=False)
    d_c, j_c = sh.dp_j(dm + 2.0 * dm2, reduce=False)
    assert float((d_c - (d_a + 2.0 * d_b)).abs().max()) <= 1e-12 * float(d_c.abs().max())
    assert float((j_c - (j_a + 2.0 * j_b)).abs().max()) <= 1e-12 * float(j_c.abs().max())
    d_ref = sh.dp(dm); j_ref = sh.j(d_ref, reduce=False)
    assert float((j_a - j_ref).abs().max()) <= 1e-13 * float(j_ref.abs().max())
