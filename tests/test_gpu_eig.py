"""GPU parity for the eigen-solvers: the one-sided Jacobi kernels behind the reference's LAPACK
wrapper names vs the same LAPACK routines the reference calls (dsyev / dspevx / dspgvx of the oracle's OpenBLAS, with the
reference's arguments).  Eigenvalues: 1e-10 relative to the spectral norm (measured ~1e-14).  Eigenvectors are defined up
to sign (and rotations inside degenerate spaces), so they are checked through sign-aligned comparison for the
non-degenerate synthetic matrices plus the invariants: residual, orthonormality (B-orthonormality for dspgvx)."""
import numpy as np
import pytest
import torch

from conftest import assert_close_1e10

pytestmark = pytest.mark.gpu


def _sym(oracle, n, seed):
    a = oracle.fill_linear(n * n, seed).reshape((n, n), order="F")
    return a + a.T


def _spd(oracle, n, seed):
    a = oracle.fill_linear(n * n, seed).reshape((n, n), order="F")
    return a @ a.T / n + np.eye(n)


def _pack(m):
    n = m.shape[0]
    return np.ascontiguousarray(np.concatenate([m[: j + 1, j] for j in range(n)]))


def _flat(m):
    return np.ascontiguousarray(m.reshape(-1, order="F"))


def _align(z, zref):
    """flip the sign of each column of z to match zref"""
    s = np.sign(np.sum(z * zref, axis=0))
    s[s == 0] = 1.0
    return z * s


def test_golden_dsyev_doc_test(rt, golden):
    """GV8: the reference's own _dsyev doc-test (matrix_blas_lapack.rs:285-317) through the drop-in names, with the
    reference's tolerance (sum of squared differences < 10E-7).  Eigenvalues as asserted there; eigenvectors column by
    column up to the sign LAPACK happens to return (ours: largest component positive)."""
    g = golden["GV8"]
    matr_b = rt.MatrixUpper.from_vec(6, np.array(g["packed"])).to_matrixfull()
    vec, w, ndim = rt._dsyev(matr_b, "V")
    assert ndim == 3
    assert np.sum((w - np.array(g["eigenvalues"])) ** 2) < g["tolerance_sum_sq"]
    assert np.max(np.abs(w - np.array(g["eigenvalues"]))) < 1e-13
    z = vec.data.reshape((3, 3), order="F"); zr = np.array(g["eigenvectors"]).reshape((3, 3), order="F")
    diff = sum(min(np.sum((z[:, c] - zr[:, c]) ** 2), np.sum((z[:, c] + zr[:, c]) ** 2)) for c in range(3))
    assert diff < g["tolerance_sum_sq"] and diff < 1e-24
    vec0, w0, _ = rt._dsyev(matr_b, "N")
    assert vec0 is None and np.sum((w0 - np.array(g["eigenvalues"])) ** 2) < g["tolerance_sum_sq"]
    # the packed form of the same matrix through lapack_dspevx
    z2, w2, found = rt.MatrixUpper.from_vec(6, np.array(g["packed"])).lapack_dspevx()
    assert found == 3 and np.max(np.abs(w2 - np.array(g["eigenvalues"]))) < 1e-13


@pytest.mark.parametrize("n", [1, 2, 3, 16, 17, 64, 257, 600])
def test_dsyev_vs_lapack(rt, oracle_blas, n):
    a = _sym(oracle_blas, n, 71)
    zref, wref = oracle_blas.dsyev(_flat(a), n)
    vec, w, nn = rt._dsyev(rt.MatrixFull.from_vec([n, n], _flat(a)), "V")
    assert nn == n and vec.size == [n, n]
    scale = max(np.max(np.abs(wref)), 1e-300)
    assert np.max(np.abs(w - wref)) <= 1e-10 * scale, np.max(np.abs(w - wref)) / scale
    assert np.all(np.diff(w) >= 0)
    z = vec.data.reshape((n, n), order="F")
    assert np.max(np.abs(a @ z - z * w)) <= 1e-11 * scale * max(n, 1)
    assert np.max(np.abs(z.T @ z - np.eye(n))) <= 1e-12 * max(n, 1)
    zr = zref.reshape((n, n), order="F")
    gaps = np.min(np.abs(np.subtract.outer(wref, wref)) + np.eye(n) * 1e300, axis=1)
    ok = gaps > 1e-6 * scale          # sign-aligned vectors agree where the eigenvalue is isolated
    assert np.max(np.abs(_align(z, zr) - zr)[:, ok]) <= 1e-8
    # largest component of every vector is positive (the documented sign convention)
    assert np.all(z[np.argmax(np.abs(z), axis=0), np.arange(n)] > 0)
    # values only
    vec0, w0, _ = rt._dsyev(rt.MatrixFull.from_vec([n, n], _flat(a)), "N")
    assert vec0 is None and np.array_equal(w0, w)


def test_dsyev_reads_the_lower_triangle_only(rt, oracle_blas):
    n = 40
    a = _sym(oracle_blas, n, 72)
    junk = a.copy()
    junk[np.triu_indices(n, 1)] = 1e30      # dsyev(.., 'L', ..): the strict upper triangle is never referenced
    _, w, _ = rt._dsyev(rt.MatrixFull.from_vec([n, n], _flat(junk)), "N")
    _, wref = oracle_blas.dsyev(_flat(a), n, "N")
    assert np.max(np.abs(w - wref)) <= 1e-10 * np.max(np.abs(wref))


def test_degenerate_and_structured_spectra(rt, oracle_blas):
    n = 48
    # exactly degenerate spectrum: Q diag(1,1,1,2,2,...) Q^T
    q, _ = np.linalg.qr(oracle_blas.fill_linear(n * n, 73).reshape((n, n), order="F"))
    lam = np.repeat(np.arange(1.0, 1.0 + n // 4), 4)
    a = (q * lam) @ q.T
    a = 0.5 * (a + a.T)
    vec, w, _ = rt._dsyev(rt.MatrixFull.from_vec([n, n], _flat(a)), "V")
    z = vec.data.reshape((n, n), order="F")
    assert np.max(np.abs(w - lam)) <= 1e-11 * n
    assert np.max(np.abs(a @ z - z * w)) <= 1e-11 * n and np.max(np.abs(z.T @ z - np.eye(n))) <= 1e-12 * n
    # diagonal, zero and identity matrices
    d = np.diag(np.arange(n, 0, -1.0))
    _, w, _ = rt._dsyev(rt.MatrixFull.from_vec([n, n], _flat(d)), "N")
    assert np.array_equal(w, np.arange(1.0, n + 1))
    _, w, _ = rt._dsyev(rt.MatrixFull.new([n, n], 0.0), "N")
    assert np.all(w == 0.0)
    with pytest.raises(rt.RestB200Error):
        rt._dsyev(rt.MatrixFull.new([3, 4], 0.0), "V")


@pytest.mark.parametrize("n", [1, 5, 64, 300])
def test_dspevx_vs_lapack(rt, oracle_blas, n):
    a = _sym(oracle_blas, n, 74)
    ap = _pack(a)
    zref, wref, mref = oracle_blas.dspevx(ap, n)
    z, w, found = rt.MatrixUpper.from_vec(ap.size, ap).lapack_dspevx()
    assert found == mref == n
    scale = np.max(np.abs(wref))
    assert np.max(np.abs(w - wref)) <= 1e-10 * scale
    zm = z.data.reshape((n, n), order="F"); zr = zref.reshape((n, n), order="F")
    assert np.max(np.abs(a @ zm - zm * w)) <= 1e-11 * scale * n
    assert np.max(np.abs(_align(zm, zr) - zr)) <= 1e-8


@pytest.mark.parametrize("n,m", [(1, 1), (6, 3), (64, 64), (264, 21), (300, 150)])
def test_dspgvx_vs_lapack(rt, oracle_blas, n, m):
    a = _sym(oracle_blas, n, 75)
    b = _spd(oracle_blas, n, 76)
    ap, bp = _pack(a), _pack(b)
    zref, wref = oracle_blas.dspgvx(ap, bp, n, m)
    z, w = rt.MatrixUpper.from_vec(ap.size, ap).lapack_dspgvx(rt.MatrixUpper.from_vec(bp.size, bp), m)
    assert z.size == [n, m] and w.size == m
    scale = np.max(np.abs(wref))
    assert np.max(np.abs(w - wref)) <= 1e-10 * scale
    zm = z.data.reshape((n, m), order="F"); zr = zref.reshape((n, m), order="F")
    assert np.max(np.abs(a @ zm - (b @ zm) * w)) <= 1e-10 * scale * n            # A z = lambda B z
    assert np.max(np.abs(zm.T @ b @ zm - np.eye(m))) <= 1e-11 * n                # B-orthonormal
    assert np.max(np.abs(_align(zm, zr) - zr)) <= 1e-7
    z2, w2 = rt._dspgvx(rt.MatrixUpper.from_vec(ap.size, ap), rt.MatrixUpper.from_vec(bp.size, bp), m)
    assert np.array_equal(w2, w) and np.array_equal(z2.data, z.data)             # deterministic


def test_dspgvx_rejects_an_indefinite_overlap(rt, oracle_blas):
    n = 12
    a = _sym(oracle_blas, n, 77)
    b = _sym(oracle_blas, n, 78)              # indefinite
    with pytest.raises(rt.RestB200Error, match="positive definite"):
        rt._dspgvx(rt.MatrixUpper.from_vec(n * (n + 1) // 2, _pack(a)), rt.MatrixUpper.from_vec(n * (n + 1) // 2, _pack(b)), 3)


@pytest.mark.parametrize("n,p", [(1, -0.5), (33, -0.5), (200, -0.5), (200, 0.5), (64, -1.0)])
def test_power_vs_reference_algorithm(rt, oracle_blas, n, p):
    s = _spd(oracle_blas, n, 79)
    ref, kept_ref = oracle_blas.power(_flat(s), n, p, 1e-10)
    got = rt._power(rt.MatrixFull.from_vec([n, n], _flat(s)), p, 1e-10)
    assert kept_ref == n
    assert_close_1e10(got.data, ref, f"_power n={n} p={p}")
    g = got.data.reshape((n, n), order="F")
    assert np.array_equal(g, g.T)
    if p == -0.5:                              # S^-1/2 S S^-1/2 = I
        assert np.max(np.abs(g @ s @ g - np.eye(n))) <= 1e-11 * n
    assert rt.MatrixFull.from_vec([n, n], _flat(s)).lapack_power(p, 1e-10).data.tobytes() == got.data.tobytes()


def test_power_drops_eigenvalues_below_the_threshold(rt, oracle_blas):
    n, r = 40, 25
    x = oracle_blas.fill_linear(n * r, 80).reshape((n, r), order="F")
    s = x @ x.T                                # rank 25: 15 (numerically) zero eigenvalues
    ref, kept_ref = oracle_blas.power(_flat(s), n, -0.5, 1e-8)
    got = rt._power(rt.MatrixFull.from_vec([n, n], _flat(s)), -0.5, 1e-8)
    assert kept_ref == r
    assert_close_1e10(got.data, ref, "pseudo-inverse square root")


def test_power_and_overlap_check_on_indefinite_input_with_symmetric_spectrum(rt, oracle_blas):
    """Matrices whose spectrum is symmetric about zero ([[0, B], [B^T, 0]]: eigenvalues +-sigma_i) are the hard case of
    the unshifted semi-definite fast path: one-sided Jacobi sees |lambda| only and the +s / -s eigenvectors can stay mixed
    (for [[0,1],[1,0]] the columns are orthogonal from the start, every Rayleigh quotient is 0).  The result must be the
    reference's: _power drops the negative eigenvalues (dsyev + threshold, matrix_blas_lapack.rs:2123-2185), and a
    non-positive-definite overlap is rejected by _dspgvx (LAPACK's Cholesky panics there)."""
    s2 = np.array([[0.0, 1.0], [1.0, 0.0]])
    got = rt._power(rt.MatrixFull.from_vec([2, 2], _flat(s2)), 1.0, 1e-10)
    assert_close_1e10(got.data, _flat(0.5 * np.ones((2, 2))), "_power([[0,1],[1,0]], p = 1)")
    for n, m in [(6, 6), (40, 23), (130, 130)]:
        b = oracle_blas.fill_linear(n * m, 81).reshape((n, m), order="F")
        s = np.block([[np.zeros((n, n)), b], [b.T, np.zeros((m, m))]])
        nn = n + m
        ref, kept_ref = oracle_blas.power(_flat(s), nn, 0.5, 1e-10)
        got = rt._power(rt.MatrixFull.from_vec([nn, nn], _flat(s)), 0.5, 1e-10)
        assert kept_ref == min(n, m)
        assert_close_1e10(got.data, ref, f"_power of the bipartite matrix {n}+{m}")
        a = _sym(oracle_blas, nn, 82)
        with pytest.raises(rt.RestB200Error, match="positive definite"):
            rt._dspgvx(rt.MatrixUpper.from_vec(nn * (nn + 1) // 2, _pack(a)), rt.MatrixUpper.from_vec(nn * (nn + 1) // 2, _pack(s)), 2)


def test_device_api_on_resident_buffers(ctx, oracle_blas):
    n = 96
    a = _sym(oracle_blas, n, 81)
    ad = torch.from_numpy(_flat(a)).to(f"cuda:{ctx.device}")
    w = ctx.empty(n); z = ctx.empty(n * n)
    ctx.dsyev("V", "U", n, ad, n, w, z, n)
    _, wref = oracle_blas.dsyev(_flat(a), n, "N")
    assert np.max(np.abs(w.cpu().numpy() - wref)) <= 1e-10 * np.max(np.abs(wref))
    out = ctx.empty(n * n)
    s = _spd(oracle_blas, n, 82)
    kept = ctx.matrix_power(n, torch.from_numpy(_flat(s)).to(f"cuda:{ctx.device}"), n, -0.5, 1e-10, out, n)
    assert kept == n
    assert_close_1e10(out.cpu().numpy(), oracle_blas.power(_flat(s), n, -0.5, 1e-10)[0], "rb_matrix_power")


def test_dsyev_beyond_the_shared_memory_cache(ctx, oracle_blas):
    """n > 3072: the column pair no longer fits the 48 KB cache and the rotation re-reads the columns through L2."""
    n = 3100
    a = _sym(oracle_blas, n, 83)
    _, wref = oracle_blas.dsyev(_flat(a), n, "N")
    ad = torch.from_numpy(_flat(a)).to(f"cuda:{ctx.device}")
    w = ctx.empty(n); z = ctx.empty(n * n)
    ctx.dsyev("V", "L", n, ad, n, w, z, n)
    wg = w.cpu().numpy()
    scale = np.max(np.abs(wref))
    assert np.max(np.abs(wg - wref)) <= 1e-10 * scale
    zt = z.view(n, n)                               # row-major view of the column-major matrix = Z^T
    at = ad.view(n, n)                              # symmetric
    res = torch.max(torch.abs(zt @ at - w[:, None] * zt)).item()
    orth = torch.max(torch.abs(zt @ zt.T - torch.eye(n, dtype=torch.float64, device=zt.device))).item()
    assert res <= 1e-11 * scale * n and orth <= 1e-12 * n, (res, orth)


def test_clustered_and_ill_conditioned_spectra(rt, oracle_blas):
    n = 120
    q, _ = np.linalg.qr(oracle_blas.fill_linear(n * n, 84).reshape((n, n), order="F"))
    # tight cluster next to isolated eigenvalues, both signs
    lam = np.concatenate([np.full(40, 1.0) + 1e-9 * np.arange(40), -np.linspace(0.5, 3.0, 40), np.linspace(10.0, 1e3, 40)])
    a = (q * lam) @ q.T
    a = 0.5 * (a + a.T)
    vec, w, _ = rt._dsyev(rt.MatrixFull.from_vec([n, n], _flat(a)), "V")
    _, wref = oracle_blas.dsyev(_flat(a), n, "N")
    z = vec.data.reshape((n, n), order="F")
    assert np.max(np.abs(w - wref)) <= 1e-10 * 1e3
    assert np.max(np.abs(a @ z - z * w)) <= 1e-10 * 1e3 and np.max(np.abs(z.T @ z - np.eye(n))) <= 1e-12 * n
    # overlap-like SPD matrix with condition 1e8: S^-1/2 S S^-1/2 = I holds to cond * eps, small eigenvalues keep their
    # relative accuracy (no shift on the semi-definite path)
    lam = np.logspace(-8, 0, n)
    s = (q * lam) @ q.T
    s = 0.5 * (s + s.T)
    x = rt._power(rt.MatrixFull.from_vec([n, n], _flat(s)), -0.5, 1e-12).data.reshape((n, n), order="F")
    assert np.max(np.abs(x @ s @ x - np.eye(n))) <= 1e-6
    vec, w, _ = rt._dsyev(rt.MatrixFull.from_vec([n, n], _flat(s)), "N")
    assert np.max(np.abs(w - lam)) <= 1e-12          # absolute accuracy relative to the norm, like LAPACK
