"""The oracle pinned against the reference's own asserted vectors (GV1-GV7) and our derived ones (GV9, GV10),
plus internal consistency: netlib-loop BLAS vs OpenBLAS, ri_ao2mo_f (restmatr.f90) vs ao2mo_v01 (pure-Rust order)."""
import numpy as np
import pytest

from conftest import assert_close_1e10


def _gemm_case(o, g):
    a = np.array(g["a"]); b = np.array(g["b"])
    c = np.full(g["size_c"][0] * g["size_c"][1], g["c_fill"])
    o.general_dgemm_f(a, g["size_a"], g["sub_a"][0], g["sub_a"][1], g["opa"], b, g["size_b"], g["sub_b"][0],
                      g["sub_b"][1], g["opb"], c, g["size_c"], g["sub_c"][0], g["sub_c"][1], g["alpha"], g["beta"])
    cm = c.reshape(g["size_c"], order="F")
    r, cc = g["sub_c"]
    blk = cm[r[0]:r[1], cc[0]:cc[1]].reshape(-1, order="F")
    outside = cm.copy()
    outside[r[0]:r[1], cc[0]:cc[1]] = g["c_fill"]
    return blk, outside


@pytest.mark.parametrize("name", ["GV1", "GV2", "GV3"])
def test_general_dgemm_f_golden(oracle, golden, name):
    g = golden[name]
    blk, outside = _gemm_case(oracle, g)
    assert blk.tolist() == g["expect_block"]
    assert np.all(outside == g["c_fill"])  # elements outside the block untouched


def test_general_dgemm_f_golden_openblas(oracle_blas, golden):
    for name in ("GV1", "GV2", "GV3"):
        blk, _ = _gemm_case(oracle_blas, golden[name])
        assert blk.tolist() == golden[name]["expect_block"]


def test_pack_order_gv4(oracle, golden):
    g = golden["GV4"]
    assert oracle.to_matrixupper(np.array(g["full"]), g["n"]).tolist() == g["packed"]


def test_unpack_gv5(oracle, golden):
    for case in golden["GV5"]["cases"]:
        assert oracle.to_matrixfull(np.array(case["packed"])).tolist() == case["full_colmajor"]


def test_unpack_edge_cases(oracle):
    assert oracle.to_matrixfull(np.zeros(4)) is None          # 4 is not triangular -> None
    assert oracle.to_matrixfull(np.zeros(0)).size == 0        # empty
    assert oracle.index2d(2, 1, 6) == 4 and oracle.index2d(1, 2, 6) == 4
    assert oracle.index2d(3, 3, 6) is None


def test_transpose_gv7(oracle, golden):
    g = golden["GV7"]
    r, c = g["size"]
    t = oracle.matrix_transpose(np.array(g["data"]), r, c).reshape((c, r), order="F")
    assert t[:, 2].tolist() == g["transposed_column_2"]


def test_ao2mo_gv9(oracle, golden):
    g = golden["GV9"]
    nb, _, nx = g["ri_size"]
    ri = np.full(nb * nb * nx, g["ri_fill"]); c = np.full(nb * nb, g["c_fill"])
    assert np.all(oracle.ri_ao2mo_f(c, ri, nb, nb, nx) == g["expect_every"])
    assert np.all(oracle.ao2mo_v01(c, ri, nb, nb, nx) == g["expect_every"])


def test_ri_transposes_gv10(oracle, golden):
    g = golden["GV10"]
    i, j, k = g["size"]
    d = np.array(g["data"])
    for which, name in enumerate(["jik", "jki", "kji", "ikj"]):
        assert oracle.ri_transpose(d, i, j, k, which).tolist() == [float(v) for v in g[name]], name


def test_ao2mo_two_restatements_agree(oracle, oracle_blas):
    nb, nx = 23, 7
    ri = oracle.fill_linear(nb * nb * nx, 2)
    c = oracle.fill_linear(nb * nb, 3, scale=nb ** -0.5)
    a = oracle.ri_ao2mo_f(c, ri, nb, nb, nx)
    b = oracle.ao2mo_v01(c, ri, nb, nb, nx)
    d = oracle_blas.ri_ao2mo_f(c, ri, nb, nb, nx)
    assert_close_1e10(a, b, "ri_ao2mo_f vs ao2mo_v01")
    assert_close_1e10(d, b, "ri_ao2mo_f(OpenBLAS) vs ao2mo_v01")
    # rectangular generalisation collapses to the square one
    assert np.array_equal(oracle.ri_ao2mo_rect(c, nb, c, nb, ri, nb, nx), a)


def test_ao2mo_against_einsum(oracle, oracle_blas):
    """Third, independent statement of ri_ao2mo_f (restmatr.f90:158-194): ri3mo[P,a,b] = sum_mu,nu C[mu,a] A[mu,nu,P] C[nu,b]
    written as ONE numpy.einsum -- no slab loop, no dgemm call order, no scatter.  Both oracle builds (netlib-style loops
    and OpenBLAS) and the rectangular generalisation must agree with it; symmetric and non-symmetric slabs."""
    for nb, nx, symm in [(17, 6, True), (23, 7, False), (40, 3, True)]:
        ri = oracle.fill_ri3ao_symm(nb, 0, nx) if symm else oracle.fill_linear(nb * nb * nx, 2)
        c = oracle.fill_linear(nb * nb, 3, scale=nb ** -0.5)
        A = ri.reshape((nb, nb, nx), order="F")
        Cm = c.reshape((nb, nb), order="F")
        ref = np.einsum("ma,mnp,nb->pab", Cm, A, Cm, optimize=False).reshape(-1, order="F")
        assert_close_1e10(oracle.ri_ao2mo_f(c, ri, nb, nb, nx), ref, "ri_ao2mo_f vs einsum")
        assert_close_1e10(oracle_blas.ri_ao2mo_f(c, ri, nb, nb, nx), ref, "ri_ao2mo_f(OpenBLAS) vs einsum")
        assert_close_1e10(oracle.ao2mo_v01(c, ri, nb, nb, nx), ref, "ao2mo_v01 vs einsum")
        no = 5
        cl, cr = np.ascontiguousarray(c[: nb * no]), np.ascontiguousarray(c[nb * no:])
        ref_ov = np.einsum("ma,mnp,nb->pab", Cm[:, :no], A, Cm[:, no:], optimize=False).reshape(-1, order="F")
        assert_close_1e10(oracle_blas.ri_ao2mo_rect(cl, no, cr, nb - no, ri, nb, nx), ref_ov, "ri_ao2mo_rect vs einsum")


def test_jk_against_einsum(oracle):
    nb, nx, no = 11, 5, 3
    ri = oracle.fill_ri3ao_symm(nb, 0, nx)
    c = oracle.fill_linear(nb * nb, 3, scale=nb ** -0.5)
    A = ri.reshape((nb, nb, nx), order="F")
    Co = c.reshape((nb, nb), order="F")[:, :no]
    D = 2.0 * Co @ Co.T
    d = oracle.ri_dp(ri, np.ascontiguousarray(D.reshape(-1, order="F")), nb, nx)
    assert_close_1e10(d, np.einsum("mnp,mn->p", A, D), "d_P")
    assert_close_1e10(oracle.ri_j(ri, d, nb, nx), np.einsum("mnp,p->mn", A, d).reshape(-1, order="F"), "J")
    ct = np.ascontiguousarray((Co * np.sqrt(2.0)).reshape(-1, order="F"))
    B = np.einsum("mnp,ni->mip", A, Co * np.sqrt(2.0))
    assert_close_1e10(oracle.ri_k(ri, ct, nb, no, nx), np.einsum("mip,lip->ml", B, B).reshape(-1, order="F"), "K")


def test_blas_layers_agree(oracle, oracle_blas):
    rng = np.random.default_rng(0)
    m, n, k = 13, 9, 17
    for ta in "NT":
        for tb in "NT":
            a = rng.standard_normal((m, k) if ta == "N" else (k, m)); b = rng.standard_normal((k, n) if tb == "N" else (n, k))
            af, bf = np.asfortranarray(a).reshape(-1, order="F"), np.asfortranarray(b).reshape(-1, order="F")
            c0 = rng.standard_normal(m * n)
            c1, c2 = c0.copy(), c0.copy()
            oracle.use_blas(False)
            oracle.dgemm(ta, tb, m, n, k, 0.7, af, a.shape[0], bf, b.shape[0], -0.3, c1, m)
            oracle_blas.dgemm(ta, tb, m, n, k, 0.7, af, a.shape[0], bf, b.shape[0], -0.3, c2, m)
            assert_close_1e10(c1, c2, f"dgemm {ta}{tb}")
            ref = 0.7 * (a if ta == "N" else a.T) @ (b if tb == "N" else b.T) - 0.3 * c0.reshape((m, n), order="F")
            assert_close_1e10(c1, ref.reshape(-1, order="F"), f"dgemm {ta}{tb} vs numpy")


def test_synth_generator_matches_numpy(oracle):
    # the splitmix64 finaliser of SURVEY 8(d), restated in numpy uint64 arithmetic
    idx = np.arange(1000, dtype=np.uint64)
    seed = np.uint64(3)
    with np.errstate(over="ignore"):
        z = seed + np.uint64(0x9E3779B97F4A7C15) * (idx + np.uint64(1))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)
    u = (z >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
    v = (2.0 * u - 1.0) * 0.25
    assert np.array_equal(oracle.fill_linear(1000, 3, 0, 0.25), v)
    a = oracle.fill_ri3ao_symm(6, 2, 4).reshape((6, 6, 2), order="F")
    assert np.array_equal(a, a.transpose(1, 0, 2))  # symmetric slabs
    assert a[1, 4, 1] == oracle.lib.orc_synth(1, 1 + 4 * 6 + 3 * 36, 1.0)


def test_einsum_helpers_oracle(oracle):
    """matrix_blas_lapack.rs:1273-1387 restated; the reference's own test (1388-1395) only prints:
    [3,4;2,6] (column-major [3,4,2,6]) with itself under "ip,ip->p" gives [25, 40]."""
    a = np.array([3.0, 4.0, 2.0, 6.0])
    assert oracle.einsum_02(a, a, 2, 2).tolist() == [25.0, 40.0]
    rng = np.random.default_rng(5)
    ni, nj = 13, 7
    m = rng.standard_normal((ni, nj)); b = rng.standard_normal(nj); v = rng.standard_normal(ni)
    mc = np.ascontiguousarray(m.reshape(-1, order="F"))
    assert np.array_equal(oracle.einsum_01(mc, b, ni, nj).reshape((ni, nj), order="F"), m * b[None, :])
    assert np.array_equal(oracle.einsum_03(v, b, ni, nj).reshape((ni, nj), order="F"), np.outer(v, b))
    assert np.allclose(oracle.einsum_02(mc, mc, ni, nj), np.einsum("ip,ip->p", m, m), rtol=1e-14)


def test_oracle_iajb_vs_numpy_einsum(oracle):
    """(ia|jb) blocks of the oracle against an independent numpy einsum over the ri3mo layout [P, l, r] (P fastest)."""
    np_, nl, nr = 23, 5, 7
    mo = oracle.fill_linear(np_ * nl * nr, 41)
    other = oracle.fill_linear(np_ * 4 * 6, 42)
    t = mo.reshape((np_, nl, nr), order="F"); u = other.reshape((np_, 4, 6), order="F")
    for (ba, bb, tb, nlb) in [((0, 5, 0, 7), (0, 5, 0, 7), t, nl), ((1, 3, 2, 4), (0, 5, 1, 6), t, nl),
                              ((2, 1, 6, 1), (0, 4, 0, 6), u, 4), ((0, 0, 0, 7), (0, 5, 0, 7), t, nl)]:
        xa = t[:, ba[0]:ba[0] + ba[1], ba[2]:ba[2] + ba[3]]
        xb = tb[:, bb[0]:bb[0] + bb[1], bb[2]:bb[2] + bb[3]]
        ref = np.einsum("pia,pjb->iajb", xa, xb).reshape(-1, order="F")
        flat_b = np.ascontiguousarray(tb.reshape(-1, order="F"))
        got = oracle.ri_iajb(np_, mo, nl, ba, flat_b, nlb, bb)
        assert got.shape == ref.shape
        if ref.size:
            assert np.max(np.abs(got - ref)) <= 1e-13 * max(1.0, np.max(np.abs(ref)))


def test_oracle_lapack_wrappers_vs_numpy(oracle_blas):
    """The oracle's eigen-solver entry points (the reference's LAPACK calls with its arguments) against numpy.linalg."""
    n = 9
    a = oracle_blas.fill_linear(n * n, 51).reshape((n, n), order="F"); a = a + a.T
    b = oracle_blas.fill_linear(n * n, 52).reshape((n, n), order="F"); b = b @ b.T + n * np.eye(n)
    flat = lambda m: np.ascontiguousarray(m.reshape(-1, order="F"))  # noqa: E731
    pack = lambda m: np.ascontiguousarray(np.concatenate([m[: j + 1, j] for j in range(n)]))  # noqa: E731
    wnp = np.linalg.eigvalsh(a)
    _, w = oracle_blas.dsyev(flat(a), n)
    assert np.max(np.abs(w - wnp)) <= 1e-13 * np.max(np.abs(wnp))
    z, w2, m = oracle_blas.dspevx(pack(a), n)
    assert m == n and np.max(np.abs(w2 - wnp)) <= 1e-13 * np.max(np.abs(wnp))
    zz, w3 = oracle_blas.dspgvx(pack(a), pack(b), n, 4)
    l = np.linalg.cholesky(b); li = np.linalg.inv(l)
    wgen = np.linalg.eigvalsh(li @ a @ li.T)[:4]
    assert np.max(np.abs(w3 - wgen)) <= 1e-12 * np.max(np.abs(wgen))
    zm = zz.reshape((n, 4), order="F")
    assert np.max(np.abs(zm.T @ b @ zm - np.eye(4))) <= 1e-12
    x, kept = oracle_blas.power(flat(b), n, -0.5, 1e-10)
    xm = x.reshape((n, n), order="F")
    assert kept == n and np.max(np.abs(xm @ b @ xm - np.eye(n))) <= 1e-12


def test_oracle_dsyev_reproduces_the_reference_doc_test(oracle_blas, golden):
    """GV8 (matrix_blas_lapack.rs:285-317): the reference's own asserted eigenpairs, with the reference's own tolerance
    (sum of squared differences < 10E-7) -- and the oracle's LAPACK even reproduces the signs to 1e-12."""
    g = golden["GV8"]
    full = oracle_blas.to_matrixfull(np.array(g["packed"]))
    vec, w = oracle_blas.dsyev(full, g["n"], "V")
    assert np.sum((w - np.array(g["eigenvalues"])) ** 2) < g["tolerance_sum_sq"]
    assert np.sum((vec - np.array(g["eigenvectors"])) ** 2) < g["tolerance_sum_sq"]
    assert np.max(np.abs(vec - np.array(g["eigenvectors"]))) < 1e-12
    none, w2 = oracle_blas.dsyev(full, g["n"], "N")
    assert none is None and np.sum((w2 - np.array(g["eigenvalues"])) ** 2) < g["tolerance_sum_sq"]
