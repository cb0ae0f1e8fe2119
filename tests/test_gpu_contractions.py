"""GPU parity, FP64 half: the DMMA GEMM core (all op combinations, TMA and generic paths, batched, triangular,
split-K), GEMV, SYMM, and the RI contractions ao2mo / d_P / J / K -- CUDA through the C ABI vs the CPU oracle
(reference algorithm: restmatr.f90 loop structure + OpenBLAS) on identical synthetic inputs.
Tolerance: 1e-10 relative (north star), norm-wise and element-wise (conftest.assert_close_1e10)."""
import numpy as np
import pytest
import torch

from conftest import assert_close_1e10, rel_err

pytestmark = pytest.mark.gpu


def _dev(ctx, a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(f"cuda:{ctx.device}")


def _col(m):
    return np.ascontiguousarray(np.asarray(m).reshape(-1, order="F"))


# ---------------------------------------------------------------- golden vectors through the compat symbols ----
@pytest.mark.parametrize("name", ["GV1", "GV2", "GV3"])
def test_golden_sub_block_gemm(rt, golden, name):
    g = golden[name]
    a = rt.MatrixFull.from_vec(g["size_a"], g["a"]); b = rt.MatrixFull.from_vec(g["size_b"], g["b"])
    c = rt.MatrixFull.new(g["size_c"], g["c_fill"])
    rt._dgemm(a, tuple(map(tuple, g["sub_a"])), g["opa"], b, tuple(map(tuple, g["sub_b"])), g["opb"], c,
              tuple(map(tuple, g["sub_c"])), g["alpha"], g["beta"])
    cm = c.data.reshape(g["size_c"], order="F")
    r, cc = g["sub_c"]
    assert cm[r[0]:r[1], cc[0]:cc[1]].reshape(-1, order="F").tolist() == g["expect_block"]
    cm[r[0]:r[1], cc[0]:cc[1]] = g["c_fill"]
    assert np.all(cm == g["c_fill"])


def test_golden_ao2mo_bench_case(rt, golden):
    g = golden["GV9"]  # benches/bench_tensors.rs: every element 200.0 exactly
    out = rt.RIFull.new(g["ri_size"], g["ri_fill"]).ao2mo_v02(rt.MatrixFull.new(g["c_size"], g["c_fill"]))
    assert out.size == [20, 10, 10]
    assert np.all(out.data == g["expect_every"])
    out = rt.RIFull.new(g["ri_size"], g["ri_fill"]).ao2mo(rt.MatrixFull.new(g["c_size"], g["c_fill"]))
    assert np.all(out.data == g["expect_every"])


# ---------------------------------------------------------------- GEMM core ----
SHAPES = [(1, 1, 1), (8, 8, 4), (17, 13, 29), (128, 128, 32), (130, 126, 70), (256, 384, 200), (300, 20, 1000),
          (64, 64, 2048), (513, 257, 129)]


@pytest.mark.parametrize("ta", "NT")
@pytest.mark.parametrize("tb", "NT")
@pytest.mark.parametrize("even", [True, False])
def test_dgemm_all_ops(ctx, oracle_blas, ta, tb, even):
    """even leading dimensions take the TMA kernel, odd ones the generic kernel; both must agree with OpenBLAS."""
    for (m, n, k) in SHAPES:
        ra, ca = (m, k) if ta == "N" else (k, m)
        rb, cb = (k, n) if tb == "N" else (n, k)

        def ld(rows):  # smallest leading dimension >= rows with the requested parity
            return rows + (rows % 2 if even else 1 - rows % 2)
        lda, ldb, ldc = ld(ra), ld(rb), ld(m)
        a = oracle_blas.fill_linear(lda * ca, 21); b = oracle_blas.fill_linear(ldb * cb, 22)
        c0 = oracle_blas.fill_linear(ldc * n, 23)
        for alpha, beta in [(1.0, 0.0), (0.7, -0.3)]:
            c_ref = c0.copy()
            oracle_blas.dgemm(ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c_ref, ldc)
            cd = _dev(ctx, c0)
            ctx.dgemm(ta, tb, m, n, k, alpha, _dev(ctx, a), lda, _dev(ctx, b), ldb, beta, cd, ldc)
            got = cd.cpu().numpy()
            assert_close_1e10(got, c_ref, f"dgemm {ta}{tb} {m}x{n}x{k} even={even}")
            # padding rows of C (between m and ldc) untouched
            if ldc > m:
                assert np.array_equal(got.reshape((ldc, n), order="F")[m:], c0.reshape((ldc, n), order="F")[m:])


def test_dgemm_generic_path_forced(ctx, oracle_blas):
    m, n, k = 200, 136, 264
    a = oracle_blas.fill_linear(m * k, 24); b = oracle_blas.fill_linear(k * n, 25)
    c_ref = np.zeros(m * n)
    oracle_blas.dgemm("N", "N", m, n, k, 1.0, a, m, b, k, 0.0, c_ref, m)
    ctx.set_gemm_path(1)
    try:
        cd = ctx.empty(m * n)
        ctx.dgemm("N", "N", m, n, k, 1.0, _dev(ctx, a), m, _dev(ctx, b), k, 0.0, cd, m)
        assert_close_1e10(cd.cpu().numpy(), c_ref, "generic kernel")
    finally:
        ctx.set_gemm_path(0)


def test_dgemm_degenerate(ctx):
    c = torch.full((12,), 3.0, dtype=torch.float64, device=f"cuda:{ctx.device}")
    a = torch.ones(12, dtype=torch.float64, device=c.device)
    ctx.dgemm("N", "N", 3, 4, 0, 1.0, a, 3, a, 1, 0.5, c, 3)      # k == 0: C = beta*C
    assert torch.all(c == 1.5)
    ctx.dgemm("N", "N", 3, 4, 2, 0.0, a, 3, a, 2, 0.0, c, 3)      # alpha == 0, beta == 0: C = 0
    assert torch.all(c == 0.0)
    ctx.dgemm("N", "N", 0, 4, 2, 1.0, a, 1, a, 2, 0.0, c, 1)      # m == 0: no-op
    with pytest.raises(Exception):
        ctx.dgemm("X", "N", 3, 4, 2, 1.0, a, 3, a, 2, 0.0, c, 3)


def test_dgemm_strided_batched_and_broadcast(ctx, oracle_blas):
    m, n, k, batch = 70, 52, 96, 5
    a = oracle_blas.fill_linear(m * k * batch, 26); b = oracle_blas.fill_linear(k * n, 27)
    c_ref = np.zeros(m * n * batch)
    for i in range(batch):
        ci = np.zeros(m * n)
        oracle_blas.dgemm("N", "N", m, n, k, 1.0, a[i * m * k:(i + 1) * m * k], m, b, k, 0.0, ci, m)
        c_ref[i * m * n:(i + 1) * m * n] = ci
    cd = ctx.empty(m * n * batch)
    ctx.dgemm_strided_batched("N", "N", m, n, k, 1.0, _dev(ctx, a), m, m * k, _dev(ctx, b), k, 0, 0.0, cd, m, m * n, batch)
    assert_close_1e10(cd.cpu().numpy(), c_ref, "batched NN with shared B")


@pytest.mark.parametrize("ta", "NT")
def test_dgemm_batched_packed_m_sweep(ctx, oracle_blas, ta):
    """Batched products with a shared B whose M is not a multiple of 128 run with "packed M" when A is MN-major ('N'): the
    16-row blocks of all batches are enumerated together, tiles straddle batches (m = 40: three batches per tile), only a
    batch's last block is ragged.  Shapes around every boundary (m < 16, m = 16 q, q 16 + 1, > 128), shallow and deep k
    (split-K partials of packed tiles), n with a ragged last column block, C with batch padding, beta != 0; 'T' runs
    the unpacked path on the same data.  Bit-identical between a batched call and per-batch calls (same tiles' sums)."""
    rng = np.random.default_rng(4242 + ord(ta))
    shapes = [(40, 60, 64, 7), (600, 60, 600, 5), (264, 21, 264, 9), (17, 130, 40, 11), (8, 8, 8, 3), (129, 5, 2100, 4),
              (250, 140, 6000, 2), (16, 64, 32, 20), (1800, 180, 96, 2)]
    for m, n, k, batch in shapes:
        lda = (m if ta == "N" else k) + 2 * int(rng.integers(0, 2))
        cols_a = k if ta == "N" else m
        sa = lda * cols_a + 2 * int(rng.integers(0, 3))
        ldc = m + 2 * int(rng.integers(0, 2)); sc = ldc * n + 2 * int(rng.integers(0, 3))
        alpha, beta = [(1.0, 0.0), (0.7, 1.0), (1.0, -0.5)][int(rng.integers(0, 3))]
        a = rng.standard_normal(sa * batch); b = rng.standard_normal(k * n); c0 = rng.standard_normal(sc * batch)
        c_ref = c0.copy()
        for i in range(batch):
            ci = c_ref[i * sc:(i + 1) * sc]
            oracle_blas.dgemm(ta, "N", m, n, k, alpha, a[i * sa:(i + 1) * sa], lda, b, k, beta, ci, ldc)
        cd = _dev(ctx, c0)
        ad, bd = _dev(ctx, a), _dev(ctx, b)
        ctx.dgemm_strided_batched(ta, "N", m, n, k, alpha, ad, lda, sa, bd, k, 0, beta, cd, ldc, sc, batch)
        what = f"batched {ta}N m={m} n={n} k={k} batch={batch} lda={lda} ldc={ldc} alpha={alpha} beta={beta}"
        got = cd.cpu().numpy()
        view = lambda x: np.stack([x[i * sc:(i + 1) * sc][: ldc * n].reshape((ldc, n), order="F") for i in range(batch)])  # noqa: E731
        assert_close_1e10(view(got)[:, :m], view(c_ref)[:, :m], what)
        assert np.array_equal(view(got)[:, m:], view(c0)[:, m:]), what + " (padding rows touched)"
        tails = lambda x: np.concatenate([x[i * sc + ldc * n:(i + 1) * sc] for i in range(batch)])  # noqa: E731
        assert np.array_equal(tails(got), tails(c0)), what + " (batch padding touched)"


def test_triangular_products_block_list_mode(ctx, oracle_blas):
    """Diagonal tiles of a triangular product run in block-list mode (36 blocks dealt to 8 warps; one operand tile when A
    and B are the same matrix).  SYRK n from one ragged block to several tiles, both triangles, both transposes, split-K,
    beta != 0, plus a triangular product of two DIFFERENT matrices with a symmetric result (the weighted RPA form)."""
    rng = np.random.default_rng(99)
    for n in [1, 15, 16, 17, 100, 128, 129, 250, 264, 383, 600, 641]:
        for trans in "NT":
            for uplo in "UL":
                k = int(rng.choice([24, 200, 5000]))
                ra, ca = (n, k) if trans == "N" else (k, n)
                lda = ra + (ra & 1); ldc = n + (n & 1)
                beta = float(rng.choice([0.0, 1.0, -0.3]))
                a = rng.standard_normal(lda * ca); c0 = rng.standard_normal(ldc * n)
                c_ref = c0.copy()
                oracle_blas.dsyrk(uplo, trans, n, k, 1.0, a, lda, beta, c_ref, ldc)
                cd = _dev(ctx, c0)
                ctx.dsyrk(uplo, trans, n, k, 1.0, _dev(ctx, a), lda, beta, cd, ldc)
                got = cd.cpu().numpy().reshape((ldc, n), order="F"); ref = c_ref.reshape((ldc, n), order="F")
                mask = (np.triu if uplo == "U" else np.tril)(np.ones((n, n), dtype=bool))
                what = f"dsyrk {uplo}{trans} n={n} k={k} beta={beta}"
                assert_close_1e10(got[:n][mask], ref[:n][mask], what)
                assert np.array_equal(got[:n][~mask], c0.reshape((ldc, n), order="F")[:n][~mask]), what + " (other triangle)"
    # weighted symmetric product through the RPA-type consumer: A = X diag(w) (a scaled copy), B = X -> two operands
    nx, nl, nr = 300, 6, 40
    mo = rng.standard_normal(nx * nl * nr); w = rng.standard_normal(nl * nr)
    out = ctx.empty(nx * nx)
    ctx.ri_mo_pq(_dev(ctx, mo), nx, nx, _dev(ctx, mo), nx, nx, nl, nr, (0, nl, 0, nr), _dev(ctx, w), 0.0, out, nx)
    x = mo.reshape((nx, nl * nr), order="F")
    assert_close_1e10(out.cpu().numpy().reshape((nx, nx), order="F"), (x * w) @ x.T, "weighted mo_pq (A != B, tri)")


def test_host_dgemm_wrappers(rt, oracle_blas):
    m, n, k = 90, 41, 77
    a = oracle_blas.fill_linear(m * k, 28); b = oracle_blas.fill_linear(n * k, 29)
    A = rt.MatrixFull.from_vec([m, k], a); B = rt.MatrixFull.from_vec([n, k], b)
    C = rt.MatrixFull.new([m, n], 1.0)
    c_ref = np.ones(m * n)
    oracle_blas.dgemm("N", "T", m, n, k, 2.0, a, m, b, n, 0.5, c_ref, m)
    rt._dgemm_full(A, "N", B, "T", C, 2.0, 0.5)
    assert_close_1e10(C.data, c_ref, "_dgemm_full NT")
    C2 = rt._dgemm_full_new(A, "N", B, "T", 2.0, 0.0)
    assert C2.size == [m, n]
    c_ref2 = np.zeros(m * n)
    oracle_blas.dgemm("N", "T", m, n, k, 2.0, a, m, b, n, 0.0, c_ref2, m)
    assert_close_1e10(C2.data, c_ref2, "_dgemm_full_new")
    D = A.ddot(rt.MatrixFull.from_vec([k, n], oracle_blas.fill_linear(k * n, 30)))
    d_ref = np.zeros(m * n)
    oracle_blas.dgemm("N", "N", m, n, k, 1.0, a, m, oracle_blas.fill_linear(k * n, 30), k, 0.0, d_ref, m)
    assert_close_1e10(D.data, d_ref, "ddot")
    assert A.ddot(A) is None


def test_dgemm_many_tiles_dynamic_scheduler(ctx, oracle_blas):
    """More tiles than SMs (ragged edges in both directions, K tail not a multiple of 8 or 32): the tiles beyond the
    first wave come from the global work counter; repeated launches must re-arm it and reproduce the same bits."""
    for ta, tb, (m, n, k) in [("T", "N", (2100, 1420, 77)), ("N", "N", (1700, 1900, 40)), ("N", "T", (2570, 1300, 9)),
                              ("T", "T", (1330, 2050, 100))]:
        ra, ca = (m, k) if ta == "N" else (k, m)
        rb, cb = (k, n) if tb == "N" else (n, k)
        a = oracle_blas.fill_linear(ra * ca, 41); b = oracle_blas.fill_linear(rb * cb, 42)
        c_ref = np.zeros(m * n)
        oracle_blas.dgemm(ta, tb, m, n, k, 1.0, a, ra, b, rb, 0.0, c_ref, m)
        ad, bd = _dev(ctx, a), _dev(ctx, b)
        first = None
        for rep in range(4):
            cd = torch.full((m * n,), float("nan"), dtype=torch.float64, device=ad.device)
            ctx.dgemm(ta, tb, m, n, k, 1.0, ad, ra, bd, rb, 0.0, cd, m)
            if first is None:
                first = cd.clone()
                assert_close_1e10(first.cpu().numpy(), c_ref, f"many-tile dgemm {ta}{tb} {m}x{n}x{k}")
            else:
                assert torch.equal(cd, first), "GEMM result must not depend on which CTA computed a tile"


def test_dgemm_random_shape_sweep(ctx, oracle_blas):
    """Seeded sweep over ops, ragged m / n (all block-row / block-column counts of an edge tile), K tails (k mod 8, k mod
    32), even leading dimensions with padding, alpha / beta, batches with and without a shared B -- TMA path vs OpenBLAS."""
    rng = np.random.default_rng(20260)
    for case in range(60):
        ta, tb = rng.choice(["N", "T"]), rng.choice(["N", "T"])
        m = int(rng.choice([rng.integers(1, 40), 128 + 16 * rng.integers(0, 8) + rng.integers(0, 16), rng.integers(200, 700)]))
        n = int(rng.choice([rng.integers(1, 40), 128 + 16 * rng.integers(0, 8) + rng.integers(0, 16), rng.integers(200, 700)]))
        k = int(rng.choice([rng.integers(1, 9), 32 * rng.integers(1, 6) + rng.integers(0, 32), rng.integers(100, 900)]))
        batch = int(rng.choice([1, 1, 3]))
        ra, ca = (m, k) if ta == "N" else (k, m)
        rb, cb = (k, n) if tb == "N" else (n, k)
        lda, ldb, ldc = ra + ra % 2 + 2 * int(rng.integers(0, 3)), rb + rb % 2 + 2 * int(rng.integers(0, 3)), m + m % 2 + 2 * int(rng.integers(0, 3))
        sa, sc = lda * ca + 2 * int(rng.integers(0, 3)), ldc * n
        shared_b = bool(rng.integers(0, 2))
        sb = 0 if shared_b else ldb * cb
        alpha, beta = [(1.0, 0.0), (0.5, 0.0), (1.0, 1.0), (-0.7, 0.3)][int(rng.integers(0, 4))]
        a = rng.standard_normal(sa * batch); b = rng.standard_normal(ldb * cb * (1 if shared_b else batch)); c0 = rng.standard_normal(sc * batch)
        c_ref = c0.copy()
        for bi in range(batch):
            cb_ = c_ref[bi * sc:(bi + 1) * sc]
            oracle_blas.dgemm(ta, tb, m, n, k, alpha, a[bi * sa:], lda, b[bi * sb:], ldb, beta, cb_, ldc)
        cd = _dev(ctx, c0)
        ctx.dgemm_strided_batched(ta, tb, m, n, k, alpha, _dev(ctx, a), lda, sa, _dev(ctx, b), ldb, sb, beta, cd, ldc, sc, batch)
        got = cd.cpu().numpy()
        what = f"case {case}: {ta}{tb} m={m} n={n} k={k} batch={batch} ld=({lda},{ldb},{ldc}) alpha={alpha} beta={beta}"
        ok_rows = np.zeros(sc * batch, dtype=bool)
        for bi in range(batch):
            blk = ok_rows[bi * sc:(bi + 1) * sc].reshape((ldc, n), order="F")
            blk[:m, :] = True
        assert_close_1e10(got[ok_rows], c_ref[ok_rows], what)
        assert np.array_equal(got[~ok_rows], c0[~ok_rows]), what + " (padding rows of C touched)"


def test_dgemm_split_k_deterministic(ctx, oracle_blas):
    """few tiles + deep K -> split-K partials reduced in a fixed order: bitwise repeatable, 1e-10 vs OpenBLAS."""
    m, n, k = 200, 130, 20000
    a = oracle_blas.fill_linear(m * k, 43); b = oracle_blas.fill_linear(k * n, 44)
    c_ref = np.zeros(m * n)
    oracle_blas.dgemm("N", "N", m, n, k, 1.0, a, m, b, k, 0.0, c_ref, m)
    ad, bd = _dev(ctx, a), _dev(ctx, b)
    outs = []
    for _ in range(3):
        cd = ctx.empty(m * n)
        ctx.dgemm("N", "N", m, n, k, 1.0, ad, m, bd, k, 0.0, cd, m)
        outs.append(cd.clone())
    assert_close_1e10(outs[0].cpu().numpy(), c_ref, "split-K dgemm")
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[1], outs[2])


def test_hand_rolled_gemm_names_and_slice_views(rt, oracle, oracle_blas):
    """_dgemm_nn / _dgemm_tn / _dgemm_tn_v02 (matrix_blas_lapack.rs:1185-1271) and the MatrixFullSlice(Mut) methods
    against the BLAS-free oracle loops (the summation order the reference's scalar code uses) and OpenBLAS."""
    m, k, n = 37, 52, 23
    a = oracle.fill_linear(m * k, 45); b = oracle.fill_linear(k * n, 46); at = oracle.fill_linear(k * m, 47)
    A = rt.MatrixFull.from_vec([m, k], a); B = rt.MatrixFull.from_vec([k, n], b); At = rt.MatrixFull.from_vec([k, m], at)
    ref_nn = np.zeros(m * n); oracle.dgemm("N", "N", m, n, k, 1.0, a, m, b, k, 0.0, ref_nn, m)
    ref_tn = np.zeros(m * n); oracle.dgemm("T", "N", m, n, k, 1.0, at, k, b, k, 0.0, ref_tn, m)
    for fn in (rt._dgemm_nn, rt._dgemm_nn_serial):
        c = fn(A.to_matrixfullslice(), B.to_matrixfullslice())
        assert c.size == [m, n]
        assert_close_1e10(c.data, ref_nn, fn.__name__)
    for fn in (rt._dgemm_tn, rt._dgemm_tn_serial):
        assert_close_1e10(fn(At.to_matrixfullslice(), B.to_matrixfullslice()).data, ref_tn, fn.__name__)
    assert rt._dgemm_nn(rt.MatrixFull.new([0, 4], 0.0), rt.MatrixFull.new([4, 3], 0.0)).size == [0, 3]
    # _dgemm_tn_v02 writes through the x-runs of a RIFull block: here the [m, n] block at (x0, :, z0..) of a tensor
    x0, sx, sy, sz = 2, m + 5, n, 3
    ten = rt.RIFull.new([sx, sy, sz], 7.0)
    rt._dgemm_tn_v02(At.to_matrixfullslice(), B.to_matrixfullslice(), ten.get_slices_mut((x0, x0 + m), (0, n), (1, 2)))
    cube = ten.data.reshape((sx, sy, sz), order="F")
    assert_close_1e10(cube[x0:x0 + m, :, 1].reshape(-1, order="F"), ref_tn, "_dgemm_tn_v02")
    mask = np.ones_like(cube, dtype=bool); mask[x0:x0 + m, :, 1] = False
    assert np.all(cube[mask] == 7.0)
    # slice methods
    assert np.array_equal(A.to_matrixfullslice().transpose().data, oracle.matrix_transpose(a, m, k))
    assert_close_1e10(A.to_matrixfullslice().ddot(B.to_matrixfullslice()).data, ref_nn, "MatrixFullSlice::ddot")
    assert A.to_matrixfullslice().ddot(A.to_matrixfullslice()) is None
    C = rt.MatrixFull.new([m, n], 1.0)
    ref = np.ones(m * n); oracle_blas.dgemm("T", "N", m, n, k, 0.5, at, k, b, k, 2.0, ref, m)
    C.to_matrixfullslicemut().lapack_dgemm(At.to_matrixfullslice(), B.to_matrixfullslice(), 'T', 'N', 0.5, 2.0)
    assert_close_1e10(C.data, ref, "MatrixFullSliceMut::lapack_dgemm")
    # MatrixUpperSlice::to_matrixfull == MatrixUpper::to_matrixfull (bit-exact data movement)
    packed = oracle.fill_linear(21, 48)
    assert np.array_equal(rt.MatrixUpperSlice.from_vec(packed).to_matrixfull().data, oracle.to_matrixfull(packed))
    assert rt.MatrixUpperSlice.from_vec(np.zeros(4)).to_matrixfull() is None


# ---------------------------------------------------------------- SYRK / SYMM / GEMV ----
@pytest.mark.parametrize("uplo", "UL")
@pytest.mark.parametrize("trans", "NT")
def test_dsyrk(rt, oracle_blas, uplo, trans):
    for n, k in [(5, 3), (130, 70), (264, 1500), (600, 64), (1500, 24)]:
        ra, ca = (n, k) if trans == "N" else (k, n)
        a = oracle_blas.fill_linear(ra * ca, 31)
        c0 = oracle_blas.fill_linear(n * n, 32)
        c_ref = c0.copy()
        oracle_blas.dsyrk(uplo, trans, n, k, 0.9, a, ra, 0.2, c_ref, n)
        C = rt.MatrixFull.from_vec([n, n], c0.copy())
        rt._dsyrk(rt.MatrixFull.from_vec([ra, ca], a), C, uplo, trans, 0.9, 0.2)
        got = C.data.reshape((n, n), order="F"); ref = c_ref.reshape((n, n), order="F"); orig = c0.reshape((n, n), order="F")
        tri = np.triu if uplo == "U" else np.tril
        other = (lambda x: np.tril(x, -1)) if uplo == "U" else (lambda x: np.triu(x, 1))
        assert_close_1e10(tri(got), tri(ref), f"dsyrk {uplo}{trans} n={n} k={k}")
        assert np.array_equal(other(got), other(orig)), "the other triangle must stay untouched"


def test_dsyrk_random_sweep(ctx, oracle_blas):
    """Seeded sweep: both triangles / transposes, ragged n across tile boundaries, shallow and deep k (split-K), even
    and odd leading dimensions (TMA and generic kernels), beta 0 / 1 / other; the other triangle stays bit-identical."""
    rng = np.random.default_rng(777)
    for case in range(40):
        uplo, trans = rng.choice(["U", "L"]), rng.choice(["N", "T"])
        n = int(rng.choice([rng.integers(1, 30), 128 * rng.integers(1, 4) + rng.integers(0, 128), rng.integers(250, 600)]))
        k = int(rng.choice([rng.integers(1, 40), rng.integers(100, 700), rng.integers(3000, 9000)]))
        ra, ca = (n, k) if trans == "N" else (k, n)
        lda = ra + int(rng.integers(0, 4)); ldc = n + int(rng.integers(0, 4))
        alpha, beta = [(1.0, 0.0), (1.0, 1.0), (0.9, -0.4)][int(rng.integers(0, 3))]
        a = rng.standard_normal(lda * ca); c0 = rng.standard_normal(ldc * n)
        c_ref = c0.copy()
        oracle_blas.dsyrk(uplo, trans, n, k, alpha, a, lda, beta, c_ref, ldc)
        cd = _dev(ctx, c0)
        ctx.dsyrk(uplo, trans, n, k, alpha, _dev(ctx, a), lda, beta, cd, ldc)
        got = cd.cpu().numpy().reshape((ldc, n), order="F"); ref = c_ref.reshape((ldc, n), order="F"); orig = c0.reshape((ldc, n), order="F")
        tri = np.triu if uplo == "U" else np.tril
        mask = tri(np.ones((n, n), dtype=bool))
        what = f"case {case}: dsyrk {uplo}{trans} n={n} k={k} lda={lda} ldc={ldc} alpha={alpha} beta={beta}"
        assert_close_1e10(got[:n][mask], ref[:n][mask], what)
        assert np.array_equal(got[:n][~mask], orig[:n][~mask]), what + " (other triangle touched)"
        assert np.array_equal(got[n:], orig[n:]), what + " (padding rows touched)"


@pytest.mark.parametrize("side", "LR")
@pytest.mark.parametrize("uplo", "UL")
def test_dsymm(rt, oracle_blas, side, uplo):
    m, n = 66, 35
    ka = m if side == "L" else n
    a = oracle_blas.fill_linear(ka * ka, 33); b = oracle_blas.fill_linear(m * n, 34); c0 = oracle_blas.fill_linear(m * n, 35)
    c_ref = c0.copy()
    oracle_blas.dsymm(side, uplo, m, n, 1.1, a, ka, b, m, -0.4, c_ref, m)
    C = rt.MatrixFull.from_vec([m, n], c0.copy())
    rt._dsymm(rt.MatrixFull.from_vec([ka, ka], a), rt.MatrixFull.from_vec([m, n], b), C, side, uplo, 1.1, -0.4)
    assert_close_1e10(C.data, c_ref, f"dsymm {side}{uplo}")


@pytest.mark.parametrize("trans", "NT")
def test_dgemv(rt, oracle_blas, trans):
    for m, n, incx, incy in [(7, 5, 1, 1), (1000, 33, 1, 1), (129, 257, 2, 3), (10000, 40, 1, 1), (64, 3000, 1, 1),
                             (33, 9, -1, 1)]:
        a = oracle_blas.fill_linear(m * n, 36)
        lenx, leny = (n, m) if trans == "N" else (m, n)
        x = oracle_blas.fill_linear(1 + (lenx - 1) * abs(incx), 37); y0 = oracle_blas.fill_linear(1 + (leny - 1) * abs(incy), 38)
        y_ref = y0.copy()
        oracle_blas.dgemv(trans, m, n, 0.8, a, m, x, incx, 0.25, y_ref, incy)
        y = y0.copy()
        rt._dgemv(rt.MatrixFull.from_vec([m, n], a), x, y, trans, 0.8, 0.25, incx, incy)
        assert_close_1e10(y, y_ref, f"dgemv {trans} {m}x{n} inc {incx},{incy}")


# ---------------------------------------------------------------- ao2mo ----
def _inputs(o, nb, nx, no, symmetric=True):
    ri = o.fill_ri3ao_symm(nb, 0, nx) if symmetric else o.fill_linear(nb * nb * nx, 2)
    c = o.fill_linear(nb * nb, 3, scale=nb ** -0.5)
    cm = c.reshape((nb, nb), order="F")
    dm = _col(2.0 * cm[:, :no] @ cm[:, :no].T)
    ct = _col(cm[:, :no] * np.sqrt(2.0))
    return ri, c, dm, ct


@pytest.mark.parametrize("nb,nx,symm", [(10, 20, True), (24, 7, False), (100, 400, True), (33, 5, False), (128, 130, False),
                                        (264, 720, True)])
def test_ao2mo_square_vs_reference_algorithm(rt, oracle_blas, nb, nx, symm):
    """square C (the only case the Fortran defines): compat symbol ri_ao2mo_f_ vs restmatr.f90 restatement + OpenBLAS.
    (100, 400) and (264, 720) are BASELINE configs A and B in full; odd nb goes through the generic (non-TMA) kernel."""
    ri, c, _, _ = _inputs(oracle_blas, nb, nx, 1, symm)
    ref = oracle_blas.ri_ao2mo_f(c, ri, nb, nb, nx)
    got = rt.RIFull.from_vec([nb, nb, nx], ri).ao2mo(rt.MatrixFull.from_vec([nb, nb], c))
    assert got.size == [nx, nb, nb]
    assert_close_1e10(got.data, ref, f"ao2mo nb={nb} nx={nx}")


def test_ao2mo_vs_blas_free_restatement(rt, oracle):
    nb, nx = 40, 9
    ri, c, _, _ = _inputs(oracle, nb, nx, 1, False)
    ref = oracle.ao2mo_v01(c, ri, nb, nb, nx)   # fixed summation order of the pure-Rust variant (ri.rs:360-379)
    got = rt.RIFull.from_vec([nb, nb, nx], ri).ao2mo_v01(rt.MatrixFull.from_vec([nb, nb], c))
    assert_close_1e10(got.data, ref, "ao2mo vs ao2mo_v01 order")


def test_ao2mo_rect_occ_vir(rt, oracle_blas):
    nb, nx, no = 64, 50, 12
    ri, c, _, _ = _inputs(oracle_blas, nb, nx, no, False)
    cm = c.reshape((nb, nb), order="F")
    cl, cr = _col(cm[:, :no]), _col(cm[:, no:])
    ref = oracle_blas.ri_ao2mo_rect(cl, no, cr, nb - no, ri, nb, nx)
    got = rt.RIFull.from_vec([nb, nb, nx], ri).ao2mo_rect(rt.MatrixFull.from_vec([nb, no], cl), rt.MatrixFull.from_vec([nb, nb - no], cr))
    assert got.size == [nx, no, nb - no]
    assert_close_1e10(got.data, ref, "ao2mo occ-vir")


def test_ao2mo_edge_cases(rt):
    assert rt.RIFull.new([4, 4, 0], 1.0).ao2mo(rt.MatrixFull.new([4, 4], 1.0)).data.size == 0        # no slabs
    out = rt.RIFull.new([1, 1, 3], 2.0).ao2mo(rt.MatrixFull.new([1, 1], 3.0))                         # 1x1 slabs
    assert out.data.tolist() == [18.0, 18.0, 18.0]


def test_ao2mo_multi_chunk_host_pipeline(rt, oracle_blas):
    """nx > 256 exercises the 3-stream H2D | compute | D2H pipeline of ri_ao2mo_f_ (several P-chunks, ragged tail)."""
    nb, nx = 48, 700
    ri, c, _, _ = _inputs(oracle_blas, nb, nx, 1, True)
    ref = oracle_blas.ri_ao2mo_f(c, ri, nb, nb, nx)
    got = rt.RIFull.from_vec([nb, nb, nx], ri).ao2mo(rt.MatrixFull.from_vec([nb, nb], c))
    assert_close_1e10(got.data, ref, "pipelined host ao2mo")


def test_host_pass_pageable_pinned_and_unbounced_agree(rt, oracle_blas):
    """The streaming host pass must give the same bits whether the caller's buffers are pinned (straight DMA), pageable
    and bounced through pinned blocks by host threads (the default for a plain Vec<f64> / numpy array), or pageable with
    the bounce disabled (driver-staged copies).  Several chunks with a ragged tail; then the caches are trimmed."""
    import ctypes as C
    import os
    from rest_tensors_b200._lib import lib, check
    nb, nx, no = 40, 777, 5
    ri, c, dm, ct = _inputs(oracle_blas, nb, nx, no, False)
    n2 = nb * nb

    def run(pinned):
        mk = (lambda n: torch.empty(n, dtype=torch.float64, pin_memory=True)) if pinned else (lambda n: torch.empty(n, dtype=torch.float64))
        ri_h = mk(nx * n2); ri_h.copy_(torch.from_numpy(ri))
        mo_h, d_h, j_h, k_h = mk(nx * n2), mk(nx), mk(n2), mk(n2)
        c_h = torch.from_numpy(c.copy()); dm_h = torch.from_numpy(dm.copy()); ct_h = torch.from_numpy(ct.copy())
        P = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
        check(lib.rb_host_ri_ao2mo_jk(P(c_h), nb, P(c_h), nb, P(ri_h), P(mo_h), nb, nx, P(dm_h), P(ct_h), no, P(d_h), P(j_h), P(k_h)),
              "rb_host_ri_ao2mo_jk")
        return [t.clone() for t in (mo_h, d_h, j_h, k_h)]

    pinned = run(True)
    bounced = run(False)
    os.environ["REST_B200_BOUNCE"] = "0"
    try:
        direct = run(False)
    finally:
        del os.environ["REST_B200_BOUNCE"]
    for a, b, d in zip(pinned, bounced, direct):
        assert torch.equal(a, b) and torch.equal(a, d)
    assert_close_1e10(pinned[0].numpy(), oracle_blas.ri_ao2mo_f(c, ri, nb, nb, nx), "host pass ao2mo")
    assert lib.rb_host_trim() == 0
    again = run(False)          # caches rebuild after a trim
    assert torch.equal(again[0], pinned[0])


@pytest.mark.parametrize("nb,nx,no,offset", [(131, 50, 9, 0), (257, 12, 30, 0), (64, 40, 8, 1)])
def test_odd_nb_and_unaligned_tensors_take_the_padded_tma_path(ctx, oracle_blas, nb, nx, no, offset):
    """Odd basis sizes (and 8-byte-aligned device tensors) cannot be described by TMA directly: ao2mo and K copy each
    chunk into zero-padded slabs and run the same GEMMs.  Parity with the reference algorithm, square and occ-vir."""
    from rest_tensors_b200.device import ShardedRI
    ri, c, dm, ct = _inputs(oracle_blas, nb, nx, no, False)
    n2 = nb * nb
    buf = ctx.empty(nx * n2 + 1)
    buf[offset: offset + nx * n2].copy_(torch.from_numpy(ri))
    sh = ShardedRI(ctx, nb, nx, data=buf[offset: offset + nx * n2])
    cbuf = ctx.empty(n2 + 1); cbuf[offset: offset + n2].copy_(torch.from_numpy(c)); cd = cbuf[offset: offset + n2]
    assert_close_1e10(sh.ao2mo(cd, nb, cd, nb).cpu().numpy(), oracle_blas.ri_ao2mo_f(c, ri, nb, nb, nx), "padded ao2mo")
    cm = c.reshape((nb, nb), order="F")
    cl, cr = _col(cm[:, :no]), _col(cm[:, no:])
    got = sh.ao2mo(cd[: nb * no], no, cd[nb * no:], nb - no)
    assert_close_1e10(got.cpu().numpy(), oracle_blas.ri_ao2mo_rect(cl, no, cr, nb - no, ri, nb, nx), "padded occ-vir ao2mo")
    ctb = ctx.empty(nb * no + 1); ctb[offset: offset + nb * no].copy_(torch.from_numpy(ct))
    k = sh.k(ctb[offset: offset + nb * no], no, reduce=False)
    assert_close_1e10(k.cpu().numpy(), oracle_blas.ri_k(ri, ct, nb, no, nx), "padded K")
    d = sh.dp(_dev(ctx, dm))
    assert_close_1e10(d.cpu().numpy(), oracle_blas.ri_dp(ri, dm, nb, nx), "d_P on an odd / unaligned tensor")
    assert_close_1e10(sh.j(d, reduce=False).cpu().numpy(), oracle_blas.ri_j(ri, oracle_blas.ri_dp(ri, dm, nb, nx), nb, nx), "J")


def test_ao2mo_device_chunked_matches_single_pass(ctx, oracle_blas):
    """Device path: a small workspace budget forces several P-chunks (strided-batched GEMM 2, ragged last chunk); the
    single-chunk path runs GEMM 2 as one flat GEMM.  Both must agree with the reference algorithm and, since every
    output element is the same fixed-order sum either way, with each other bit for bit."""
    from rest_tensors_b200.device import ShardedRI
    nb, nx = 72, 333
    ri, c, _, _ = _inputs(oracle_blas, nb, nx, 1, False)
    ref = oracle_blas.ri_ao2mo_f(c, ri, nb, nb, nx)
    sh = ShardedRI(ctx, nb, nx, data=_dev(ctx, ri))
    cd = _dev(ctx, c)
    one = sh.ao2mo(cd, nb, cd, nb).clone()
    assert_close_1e10(one.cpu().numpy(), ref, "device ao2mo, one chunk")
    # sub-shards written into a wider output (out_ldp = nx > chunk) take the batched path
    out = ctx.empty(nx * nb * nb)
    for p0, p1 in [(0, 128), (128, 300), (300, nx)]:
        part = ShardedRI(ctx, nb, p1 - p0, data=sh.data[nb * nb * p0: nb * nb * p1].clone())
        part.ao2mo(cd, nb, cd, nb, out=out[p0:], out_ldp=nx)
    assert torch.equal(out, one), "chunked (batched GEMM 2) and single-pass (flat GEMM 2) ao2mo must agree bitwise"


def test_ao2mo_full_size_identity_property(ctx):
    """BASELINE config B size (nb=264, nx=720) on device: with C = I the transform must return the input bits,
    ri3mo[P,a,b] == ri3ao[a,b,P] (every sum has one non-zero term), and scaling C by 2 scales the result by 4 exactly."""
    nb, nx = 264, 720
    from rest_tensors_b200.device import ShardedRI
    ri = ShardedRI(ctx, nb, nx).fill_synthetic()
    eye = torch.eye(nb, dtype=torch.float64, device=ri.data.device).reshape(-1).contiguous()
    out = ri.ao2mo(eye, nb, eye, nb)
    # torch views are row-major: data.view(nx, nb(nu), nb(mu))[P, nu, mu]; out is column-major [P, a, b]
    out_rm = out.view(nb, nb, nx)                                               # [b, a, P]
    assert torch.equal(out_rm, ri.data.view(nx, nb, nb).permute(1, 2, 0)), "ao2mo(I) must reproduce ri3ao bit for bit"
    c = ctx.empty(nb * nb)
    ctx.fill_linear(c, nb * nb, 3, 0, nb ** -0.5)
    o1 = ri.ao2mo(c, nb, c, nb)
    c2 = c * 2.0
    o2 = ri.ao2mo(c2, nb, c2, nb)
    assert torch.equal(o2, o1 * 4.0)
    # symmetric slabs + same C on both sides => ri3mo[P,a,b] == ri3mo[P,b,a] up to rounding
    o1v = o1.view(nb, nb, nx)
    assert rel_err(o1v.permute(1, 0, 2).cpu().numpy(), o1v.cpu().numpy()) < 1e-12


@pytest.mark.parametrize("name,nb,nx,no", [("C", 600, 1700, 60), ("D-shard", 1800, 600, 180)])
def test_full_size_configs_properties(ctx, name, nb, nx, no):
    """BASELINE configs C (whole) and D (one rank's 600-slab shard of naux = 4800) at full size, device-resident.
    Size-independent properties instead of an oracle run: ao2mo(I) returns the input bits; ao2mo(2C) == 4 ao2mo(C)
    exactly; d_P/J linear; K symmetric with a non-negative diagonal; two half-shards add up to the whole (the
    multi-GPU invariant); occ-vir ao2mo equals the matching sub-block of the square one bit for bit."""
    from rest_tensors_b200.device import ShardedRI
    ri = ShardedRI(ctx, nb, nx).fill_synthetic()
    n2 = nb * nb
    eye = torch.eye(nb, dtype=torch.float64, device=ri.data.device).reshape(-1).contiguous()
    out = ri.ao2mo(eye, nb, eye, nb)
    src = ri.data.view(nx, nb, nb)                       # [P, nu, mu] (row-major view of the column-major tensor)
    for b0 in range(0, nb, 256):                          # compare in slices: no 15 GB temporaries
        b1 = min(nb, b0 + 256)
        assert torch.equal(out.view(nb, nb, nx)[b0:b1], src[:, b0:b1, :].permute(1, 2, 0)), f"{name}: ao2mo(I) != ri3ao"
    c = ctx.empty(n2); ctx.fill_linear(c, n2, 3, 0, nb ** -0.5)
    ri.ao2mo(c, nb, c, nb, out=out)
    probe = out[:: 1000003].clone()
    sub_sq = out.view(nb, nb, nx)[no:no + 7, :no, :].clone()           # [b in vir, a in occ, P]
    c2 = c * 2.0
    out2 = ri.ao2mo(c2, nb, c2, nb, out=out)
    assert torch.equal(out2[:: 1000003], probe * 4.0), f"{name}: power-of-two scaling must be exact"
    # occ-vir (north-star form) on a few virtual columns == sub-block of the square transform (same sums, same order)
    ov = ri.ao2mo(c[: nb * no], no, c[nb * no: nb * (no + 7)], 7)
    assert torch.equal(ov.view(7, no, nx), sub_sq), f"{name}: occ-vir block differs from the square transform"
    del out, out2, ov
    torch.cuda.empty_cache()
    cm = c.view(nb, nb).t()
    dm = (2.0 * cm[:, :no] @ cm[:, :no].t()).t().contiguous().reshape(-1)
    ct = (cm[:, :no] * (2.0 ** 0.5)).t().contiguous().reshape(-1)
    d = ri.dp(dm); j = ri.j(d, reduce=False); k = ri.k(ct, no, reduce=False)
    assert torch.equal(ri.dp(dm * 2.0), d * 2.0)
    jm, km = j.view(nb, nb), k.view(nb, nb)
    assert rel_err(jm.t().cpu().numpy(), jm.cpu().numpy()) < 1e-12
    assert torch.equal(km, km.t()) and bool((torch.diagonal(km) >= 0).all())
    half = (nx // 2) // 8 * 8 + 4                          # deliberately not a multiple of 8
    lo = ShardedRI(ctx, nb, half, data=ri.data[: n2 * half])
    hi = ShardedRI(ctx, nb, nx - half, data=ri.data[n2 * half:])
    assert rel_err(torch.cat([lo.dp(dm), hi.dp(dm)]).cpu().numpy(), d.cpu().numpy()) < 1e-12
    assert rel_err((lo.j(d[:half], reduce=False) + hi.j(d[half:], reduce=False)).cpu().numpy(), j.cpu().numpy()) < 1e-12
    assert rel_err((lo.k(ct, no, reduce=False) + hi.k(ct, no, reduce=False)).cpu().numpy(), k.cpu().numpy()) < 1e-12


def test_config_c_full_size_parity_vs_oracle(ctx, oracle_blas):
    """The bench workload itself (config C: nb=600, naux=1700, nocc=60; 4.9 GB in, 4.9 GB out) against the reference
    algorithm on the host (restmatr.f90 loop structure + OpenBLAS, ~10 s): ao2mo, d_P, J and K at full size, 1e-10."""
    import os
    try:
        avail = int([l for l in open("/proc/meminfo") if l.startswith("MemAvailable:")][0].split()[1]) * 1024
    except Exception:
        avail = None
    if avail is not None and avail < 24e9:
        pytest.skip("needs ~20 GB of host memory for the oracle's copy of config C")
    from rest_tensors_b200.device import ShardedRI
    nb, nx, no = 600, 1700, 60
    oracle_blas.set_threads(len(os.sched_getaffinity(0)))
    ri, c, dm, ct = _inputs(oracle_blas, nb, nx, no, True)
    sh = ShardedRI(ctx, nb, nx).fill_synthetic()
    assert np.array_equal(sh.data[: nb * nb * 3].cpu().numpy(), ri[: nb * nb * 3])
    cd = _dev(ctx, c)
    got = sh.ao2mo(cd, nb, cd, nb).cpu().numpy()
    ref = oracle_blas.ri_ao2mo_f(c, ri, nb, nb, nx)
    scale = float(np.max(np.abs(ref)))
    assert float(np.max(np.abs(got - ref))) / scale <= 1e-10
    assert_close_1e10(got[::97], ref[::97], "config C ao2mo (every 97th element, element-wise bound)")
    del got, ref
    d_ref = oracle_blas.ri_dp(ri, dm, nb, nx)
    d = sh.dp(_dev(ctx, dm))
    assert_close_1e10(d.cpu().numpy(), d_ref, "config C d_P")
    assert_close_1e10(sh.j(d, reduce=False).cpu().numpy(), oracle_blas.ri_j(ri, d_ref, nb, nx), "config C J")
    assert_close_1e10(sh.k(_dev(ctx, ct), no, reduce=False).cpu().numpy(), oracle_blas.ri_k(ri, ct, nb, no, nx), "config C K")


def test_config_d_shapes_subset_parity_vs_oracle(ctx, oracle_blas):
    """Config D's shapes (nb=1800, nocc=180) against the oracle on a subset of the north-star shard: per-slab work is
    independent, so slabs [1234, 1234+12) of the naux=4800 tensor exercise the same tiles / chunk pitches as the full
    600-slab shard (18 ragged k8 blocks, 15 tile columns, N = nocc = 180).  Square and occ-vir ao2mo, d_P, J, K: 1e-10."""
    import os
    from rest_tensors_b200.device import ShardedRI
    nb, no, p_lo, ns = 1800, 180, 1234, 12
    oracle_blas.set_threads(len(os.sched_getaffinity(0)))
    ri = oracle_blas.fill_ri3ao_symm(nb, p_lo, p_lo + ns)
    c = oracle_blas.fill_linear(nb * nb, 3, scale=nb ** -0.5)
    cm = c.reshape((nb, nb), order="F")
    dm = np.ascontiguousarray((2.0 * cm[:, :no] @ cm[:, :no].T).reshape(-1, order="F"))
    ct = np.ascontiguousarray((cm[:, :no] * np.sqrt(2.0)).reshape(-1, order="F"))
    sh = ShardedRI(ctx, nb, ns)
    ctx.fill_ri3ao_symm(sh.data, nb, p_lo, p_lo + ns, 1, 1.0)
    assert np.array_equal(sh.data.cpu().numpy(), ri), "device generator != oracle generator at config D offsets"
    cd = _dev(ctx, c)
    got = sh.ao2mo(cd, nb, cd, nb).cpu().numpy()
    ref = oracle_blas.ri_ao2mo_f(c, ri, nb, nb, ns)
    assert_close_1e10(got, ref, "config D shapes: square ao2mo")
    nv = nb - no
    ov = sh.ao2mo(cd[: nb * no], no, cd[nb * no:], nv).cpu().numpy()
    ov_ref = oracle_blas.ri_ao2mo_rect(np.ascontiguousarray(c[: nb * no]), no, np.ascontiguousarray(c[nb * no:]), nv, ri, nb, ns)
    assert_close_1e10(ov, ov_ref, "config D shapes: occ-vir ao2mo")
    # the occ-vir block is the [occ, vir] sub-block of the square transform
    assert_close_1e10(ov, ref.reshape((ns, nb, nb), order="F")[:, :no, no:].reshape(-1, order="F"), "occ-vir vs square block")
    d_ref = oracle_blas.ri_dp(ri, dm, nb, ns)
    d = sh.dp(_dev(ctx, dm))
    assert_close_1e10(d.cpu().numpy(), d_ref, "config D shapes: d_P")
    assert_close_1e10(sh.j(d, reduce=False).cpu().numpy(), oracle_blas.ri_j(ri, d_ref, nb, ns), "config D shapes: J")
    assert_close_1e10(sh.k(_dev(ctx, ct), no, reduce=False).cpu().numpy(), oracle_blas.ri_k(ri, ct, nb, no, ns), "config D shapes: K")


# ---------------------------------------------------------------- d_P, J, K ----
@pytest.mark.parametrize("nb,nx,no,symm", [(10, 20, 3, True), (100, 400, 20, True), (37, 11, 5, False), (64, 300, 64, False),
                                           (264, 720, 21, True)])
def test_dp_j_k_vs_oracle(rt, oracle_blas, nb, nx, no, symm):
    ri, c, dm, ct = _inputs(oracle_blas, nb, nx, no, symm)
    R = rt.RIFull.from_vec([nb, nb, nx], ri)
    d_ref = oracle_blas.ri_dp(ri, dm, nb, nx)
    d = R.ri_dp(rt.MatrixFull.from_vec([nb, nb], dm))
    assert_close_1e10(d, d_ref, "d_P")
    assert_close_1e10(R.ri_j(d_ref).data, oracle_blas.ri_j(ri, d_ref, nb, nx), "J")
    assert_close_1e10(R.ri_k(rt.MatrixFull.from_vec([nb, no], ct)).data, oracle_blas.ri_k(ri, ct, nb, no, nx), "K")


def test_jk_full_size_properties(ctx):
    """config B size on device: J symmetric for symmetric slabs, K symmetric with non-negative diagonal,
    linearity of d_P and J, and P-shard additivity (sum of two half-shards == whole) -- the multi-GPU invariant."""
    nb, nx, no = 264, 720, 21
    from rest_tensors_b200.device import ShardedRI
    ri = ShardedRI(ctx, nb, nx).fill_synthetic()
    c = ctx.empty(nb * nb); ctx.fill_linear(c, nb * nb, 3, 0, nb ** -0.5)
    cm = c.view(nb, nb).t()                      # column-major [nb, nb] as a torch matrix
    dm = (2.0 * cm[:, :no] @ cm[:, :no].t()).t().contiguous().reshape(-1)
    ct = (cm[:, :no] * (2.0 ** 0.5)).t().contiguous().reshape(-1)
    d = ri.dp(dm)
    j = ri.j(d, reduce=False)
    k = ri.k(ct, no, reduce=False)
    jm, km = j.view(nb, nb), k.view(nb, nb)
    assert rel_err(jm.t().cpu().numpy(), jm.cpu().numpy()) < 1e-12
    assert torch.equal(km, km.t()) and bool((torch.diagonal(km) >= 0).all())
    d2 = ri.dp(dm * 2.0)
    assert torch.equal(d2, d * 2.0)
    # shard additivity
    half = nx // 2
    lo = ShardedRI(ctx, nb, nx, rank=0, world=2, data=ri.data[: nb * nb * half])
    hi = ShardedRI(ctx, nb, nx, rank=1, world=2, data=ri.data[nb * nb * half:])
    assert (lo.p_lo, lo.p_hi, hi.p_lo, hi.p_hi) == (0, half, half, nx)
    assert rel_err(torch.cat([lo.dp(dm), hi.dp(dm)]).cpu().numpy(), d.cpu().numpy()) < 1e-12
    jsum = lo.j(d[:half], reduce=False) + hi.j(d[half:], reduce=False)
    ksum = lo.k(ct, no, reduce=False) + hi.k(ct, no, reduce=False)
    assert rel_err(jsum.cpu().numpy(), j.cpu().numpy()) < 1e-12
    assert rel_err(ksum.cpu().numpy(), k.cpu().numpy()) < 1e-12


def test_fused_streaming_step(rt, oracle_blas):
    """ao2mo_jk: one upload per P-chunk feeding ao2mo, d_P, J (accumulated) and K (accumulated), several chunks."""
    nb, nx, no = 40, 600, 6
    ri, c, dm, ct = _inputs(oracle_blas, nb, nx, no, True)
    mo, d, j, k = rt.RIFull.from_vec([nb, nb, nx], ri).ao2mo_jk(rt.MatrixFull.from_vec([nb, nb], c),
                                                                   rt.MatrixFull.from_vec([nb, nb], dm),
                                                                   rt.MatrixFull.from_vec([nb, no], ct))
    d_ref = oracle_blas.ri_dp(ri, dm, nb, nx)
    assert_close_1e10(mo.data, oracle_blas.ri_ao2mo_f(c, ri, nb, nb, nx), "fused ao2mo")
    assert_close_1e10(d, d_ref, "fused d_P")
    assert_close_1e10(j.data, oracle_blas.ri_j(ri, d_ref, nb, nx), "fused J")
    assert_close_1e10(k.data, oracle_blas.ri_k(ri, ct, nb, no, nx), "fused K")


@pytest.mark.parametrize("nb,ns,nx,no,pinned,symm", [(40, 40, 600, 6, False, True), (33, 21, 301, 5, False, True),
                                                         (64, 64, 520, 9, True, True), (24, 24, 50, 4, False, False)])
def test_fused_streaming_step_upper_output(rt, oracle_blas, nb, ns, nx, no, pinned, symm):
    """ao2mo_jk_upper ships only the a <= b pairs of ri3mo (MatrixUpper's pair index per P, P fastest): bit for bit the a <= b
    entries of the full pass, equal to the oracle's ri_ao2mo_f, with d_P / J / K unchanged; odd chunk sizes (nx = 301: scalar
    pack kernel), several chunks, a rectangular coefficient block (ns < nb), pageable and page-locked host buffers."""
    import ctypes as C
    from rest_tensors_b200._lib import lib, check
    ri, c_full, dm, ct = _inputs(oracle_blas, nb, nx, no, symm)
    c = np.ascontiguousarray(c_full[: nb * ns])
    if pinned:
        check(lib.rb_host_register(C.c_void_p(ri.ctypes.data), ri.nbytes), "rb_host_register")
    try:
        rif = rt.RIFull.from_vec([nb, nb, nx], ri)
        args = (rt.MatrixFull.from_vec([nb, ns], c), rt.MatrixFull.from_vec([nb, nb], dm), rt.MatrixFull.from_vec([nb, no], ct))
        upper, d, j, k = rif.ao2mo_jk_upper(*args)
        mo, d2, j2, k2 = rif.ao2mo_jk(*args)
    finally:
        if pinned:
            check(lib.rb_host_unregister(C.c_void_p(ri.ctypes.data)), "rb_host_unregister")
    full = mo.data.reshape((nx, ns, ns), order="F")
    up = upper.reshape((nx, ns * (ns + 1) // 2), order="F")
    for b in range(ns):
        for a in range(b + 1):
            assert np.array_equal(up[:, b * (b + 1) // 2 + a], full[:, a, b]), (a, b)
    assert np.array_equal(d, d2) and np.array_equal(j.data, j2.data) and np.array_equal(k.data, k2.data)
    w3 = np.asarray(oracle_blas.ri_ao2mo_rect(c, ns, c, ns, ri, nb, nx)).reshape((nx, ns, ns), order="F")
    ref_up = np.stack([w3[:, a, b] for b in range(ns) for a in range(b + 1)], axis=1)
    assert_close_1e10(up.ravel(order="F"), ref_up.ravel(order="F"), "upper ao2mo vs oracle")


@pytest.mark.parametrize("nb,ns,nx,no,pinned", [(40, 40, 600, 6, False), (33, 21, 301, 5, False), (64, 64, 520, 9, True), (100, 100, 40, 20, True),
                                                    (1, 1, 5, 1, False)])
def test_fused_streaming_step_symmetric_slabs(rt, oracle_blas, nb, ns, nx, no, pinned):
    """rb_host_ri_ao2mo_jk_symm uploads only mu <= nu of every slab (32-column trapezoids, pitched 3-D copies) and mirrors on the device:
    for symmetric slabs every output is bit for bit what the full upload gives, and the strict lower triangle of the host tensor is never
    read -- it is filled with NaN here."""
    import ctypes as C
    from rest_tensors_b200._lib import lib, check
    ri, c_full, dm, ct = _inputs(oracle_blas, nb, nx, no, True)
    c = np.ascontiguousarray(c_full[: nb * ns])
    args = (rt.MatrixFull.from_vec([nb, ns], c), rt.MatrixFull.from_vec([nb, nb], dm), rt.MatrixFull.from_vec([nb, no], ct))
    upper, d, j, k = rt.RIFull.from_vec([nb, nb, nx], ri).ao2mo_jk_upper(*args)
    holed = ri.copy().reshape((nb, nb, nx), order="F")
    holed[np.tril_indices(nb, -1)] = np.nan
    holed = np.ascontiguousarray(holed.ravel(order="F"))
    if pinned:
        check(lib.rb_host_register(C.c_void_p(holed.ctypes.data), holed.nbytes), "rb_host_register")
    try:
        upper2, d2, j2, k2 = rt.RIFull.from_vec([nb, nb, nx], holed).ao2mo_jk_upper(*args, symmetric_slabs=True)
    finally:
        if pinned:
            check(lib.rb_host_unregister(C.c_void_p(holed.ctypes.data)), "rb_host_unregister")
    assert np.array_equal(upper2, upper) and np.array_equal(d2, d)
    assert np.array_equal(j2.data, j.data) and np.array_equal(k2.data, k.data)
    assert_close_1e10(k2.data, oracle_blas.ri_k(ri, ct, nb, no, nx), "K from half-uploaded slabs")


# ---------------------------------------------------------------- special_dgemm_f_01 ----
def test_special_dgemm_f_01(rt, oracle_blas):
    X, Y, Z = 12, 7, 10
    sx, lx, sz, lz = 2, 8, 1, 6
    t0 = oracle_blas.fill_linear(X * Y * Z, 39)
    b = oracle_blas.fill_linear(9 * 9, 40)
    t_ref = t0.copy()
    oracle_blas.special_dgemm_f_01(t_ref, [X, Y, Z], (sx, sx + lx), 0, (sz, sz + lz), b, [9, 9], (1, 1 + lz), (2, 2 + lz), 0.7, 0.1)
    t = t0.copy()
    rt.special_dgemm_f_01(t, [X, Y, Z], (sx, sx + lx), 0, (sz, sz + lz), b, [9, 9], (1, 1 + lz), (2, 2 + lz), 0.7, 0.1)
    assert_close_1e10(t, t_ref, "special_dgemm_f_01")
    mask = np.ones((X, Y, Z), dtype=bool); mask[sx:sx + lx, :, sz:sz + lz] = False
    assert np.array_equal(t.reshape((X, Y, Z), order="F")[mask], t0.reshape((X, Y, Z), order="F")[mask])


def test_special_dgemm_over_p_single_rank(ctx, oracle_blas):
    """ShardedRI.special_dgemm_p at world == 1 == special_dgemm_f_01 with the full x and z ranges on the resident tensor."""
    from rest_tensors_b200.device import ShardedRI
    nb, nx = 24, 37
    sh = ShardedRI(ctx, nb, nx).fill_synthetic()
    t_ref = oracle_blas.fill_ri3ao_symm(nb, 0, nx)
    b = oracle_blas.fill_linear(nx * nx, 8, scale=nx ** -0.5)
    oracle_blas.special_dgemm_f_01(t_ref, [nb, nb, nx], (0, nb), 0, (0, nx), b, [nx, nx], (0, nx), (0, nx), 0.7, -0.2)
    sh.special_dgemm_p(_dev(ctx, b), 0.7, -0.2)
    assert_close_1e10(sh.data.cpu().numpy(), t_ref, "special_dgemm over P, one rank")


@pytest.mark.parametrize("n,k", [(1800, 20000), (1800, 4000), (1544, 20000), (1032, 20000), (776, 4000), (600, 54720)])
def test_split_k_products_repeat_bit_for_bit(ctx, n, k):
    """Regression test for a race found in round 2 (an epilogue refactor made repeated split-K SYRK / GEMM launches differ
    in single 16 x 64 strips, 1e-4 relative; it passed every parity test at the sizes the oracle runs and failed only
    the half-shard additivity property at config D): every launch of the same product must return the same bits, and the
    SYRK triangle must agree with the full product."""
    a = ctx.empty(n * k); ctx.fill_linear(a, n * k, 7, 0, 1.0)
    tri = []
    for _ in range(4):
        c = ctx.empty(n * n); c.fill_(float("nan"))
        ctx.dsyrk("U", "N", n, k, 1.0, a, n, 0.0, c, n)
        tri.append(torch.triu(c.view(n, n).t()).clone())
    for r in range(1, 4):
        assert torch.equal(tri[r], tri[0]), f"SYRK run {r} differs from run 0 in {int((tri[r] != tri[0]).sum())} elements"
    full = []
    for _ in range(3):
        c = ctx.empty(n * n); c.fill_(float("nan"))
        ctx.dgemm("N", "T", n, n, k, 1.0, a, n, a, n, 0.0, c, n)
        full.append(c.clone())
    assert torch.equal(full[1], full[0]) and torch.equal(full[2], full[0])
    ref = torch.triu(full[0].view(n, n).t())
    err = float((tri[0] - ref).abs().max() / ref.abs().max())
    assert err < 1e-12, f"SYRK triangle vs full product: {err:.3e}"


@pytest.mark.parametrize("m,n,k,batch", [(500, 500, 500, 1), (1000, 1000, 1000, 1), (2000, 2000, 2000, 1), (600, 600, 102000, 1),
                                         (264, 264, 15120, 1), (1800, 1800, 54720, 1), (300, 200, 5000, 3), (129, 1030, 777, 1),
                                         (136, 136, 40000, 1)])
def test_stream_k_products(ctx, m, n, k, batch):
    """Shapes the planner runs as stream-K (every CTA an equal share of the tile x k-step space, pieces summed in piece order by
    rb_streamk_reduce_kernel): GEMM in all four layouts with alpha / beta, SYRK, batched; against torch's FP64 matmul (cuBLAS, an
    independent implementation) to 1e-11 of the largest element (k up to 1e5 terms per element), and bit for bit across repeated launches."""
    from rest_tensors_b200._lib import lib
    assert lib.rb_gemm_plan_stream_k(m, n, k, batch, 0, ctx.num_sms, None) == 1, "this shape is meant to exercise the stream-K path"
    a = ctx.empty(batch * m * k); b = ctx.empty(batch * k * n); c0 = ctx.empty(batch * m * n)
    ctx.fill_linear(a, a.numel(), 41, 0, 1.0); ctx.fill_linear(b, b.numel(), 42, 0, 1.0); ctx.fill_linear(c0, c0.numel(), 43, 0, 1.0)
    for ta, tb in (("N", "N"), ("T", "N"), ("N", "T"), ("T", "T")):
        if batch > 1 and (ta, tb) != ("N", "N"):
            continue
        lda = m if ta == "N" else k
        ldb = k if tb == "N" else n
        outs = []
        for rep in range(2):
            c = c0.clone()
            if batch == 1:
                ctx.dgemm(ta, tb, m, n, k, 0.7, a, lda, b, ldb, -0.3, c, m)
            else:
                ctx.dgemm_strided_batched(ta, tb, m, n, k, 0.7, a, lda, m * k, b, ldb, k * n, -0.3, c, m, m * n, batch)
            outs.append(c)
        assert torch.equal(outs[0], outs[1]), (ta, tb, "repeated launches differ")
        for bi in range(batch):
            am = a[bi * m * k:(bi + 1) * m * k].view(k, m).t() if ta == "N" else a.view(m, k)      # op(A) [m, k]
            bm = b[bi * k * n:(bi + 1) * k * n].view(n, k).t() if tb == "N" else b.view(k, n)      # op(B) [k, n]
            want = 0.7 * (am @ bm) - 0.3 * c0[bi * m * n:(bi + 1) * m * n].view(n, m).t()
            got = outs[0][bi * m * n:(bi + 1) * m * n].view(n, m).t()
            err = float((got - want).abs().max() / want.abs().max())
            assert err < 1e-11, (ta, tb, bi, err)
    if m == n and batch == 1:
        for uplo in ("U", "L"):
            c = c0.clone()
            ctx.dsyrk(uplo, "N", n, k, 1.0, a, n, 0.5, c, n)
            am = a.view(k, n).t()
            want = am @ am.t() + 0.5 * c0.view(n, n).t()
            got = c.view(n, n).t()
            tri = torch.triu(torch.ones(n, n, dtype=torch.bool, device=c.device)) if uplo == "U" else \
                torch.tril(torch.ones(n, n, dtype=torch.bool, device=c.device))
            assert float((got - want)[tri].abs().max() / want.abs().max()) < 1e-11, uplo
            assert torch.equal(got[~tri], c0.view(n, n).t()[~tri]), "the other triangle must stay untouched"


def test_host_register_makes_a_caller_buffer_pinned(rt, oracle_blas):
    """rb_host_register page-locks a caller-owned (numpy) buffer in place: the streaming pass must give the same bits as with
    the pageable bounce, and the buffer must be released cleanly."""
    import ctypes as C
    from rest_tensors_b200._lib import lib, check
    nb, nx = 48, 300
    ri = oracle_blas.fill_ri3ao_symm(nb, 0, nx)
    c = oracle_blas.fill_linear(nb * nb, 3, scale=nb ** -0.5)
    ref = rt.RIFull.from_vec([nb, nb, nx], ri).ao2mo(rt.MatrixFull.from_vec([nb, nb], c)).data.copy()
    check(lib.rb_host_register(C.c_void_p(ri.ctypes.data), ri.nbytes), "rb_host_register")
    try:
        got = rt.RIFull.from_vec([nb, nb, nx], ri).ao2mo(rt.MatrixFull.from_vec([nb, nb], c)).data
        assert np.array_equal(got, ref)
    finally:
        check(lib.rb_host_unregister(C.c_void_p(ri.ctypes.data)), "rb_host_unregister")
    assert lib.rb_host_unregister(C.c_void_p(ri.ctypes.data)) != 0      # a second release is an error, not a crash
    # ... and a reported CUDA error must not linger in the runtime: the next launch check has to see a clean state
    again = rt.RIFull.from_vec([nb, nb, nx], ri).ao2mo(rt.MatrixFull.from_vec([nb, nb], c)).data
    assert np.array_equal(again, ref)
    vec, w, n3 = rt._dsyev(rt.MatrixFull.from_vec([3, 3], np.array([2.0, 1, 0, 1, 2, 1, 0, 1, 2])), "V")
    assert n3 == 3 and abs(w[1] - 2.0) < 1e-13


def test_partials_never_read_stale_workspace(ctx):
    """Every kernel that sums workspace partials (split-K and stream-K GEMM / SYRK, the packed-M half-transform of the K build, GEMV
    column chunks, ao2mo panels, the eigen-solver's scratch) must read only what the same call wrote: with every workspace poisoned
    to NaN right before the call the result has to be finite and bit-identical to the unpoisoned run.  (A stale read is how a wrong
    tile / piece bookkeeping shows up as run-to-run differences instead of a plain wrong answer.)"""
    from rest_tensors_b200.device import ShardedRI
    from rest_tensors_b200._lib import lib

    def both(fn):
        r0 = fn().clone()
        ctx.poison_workspaces()
        r1 = fn().clone()
        assert bool(torch.isfinite(r1).all()), "result depends on workspace memory the call did not write"
        assert torch.equal(r0, r1)

    kinds = set()
    for (m, n, k, batch) in [(500, 500, 500, 1), (264, 264, 15120, 1), (129, 1030, 777, 1), (300, 200, 5000, 3), (600, 600, 102000, 1),
                             (1800, 1800, 19200, 1), (200, 130, 6000, 1), (1000, 72, 9000, 1), (40, 40, 20000, 1)]:
        kinds.add(int(lib.rb_gemm_plan_stream_k(m, n, k, batch, 0, ctx.num_sms, None)))
        a = ctx.empty(batch * m * k); b = ctx.empty(batch * k * n)
        ctx.fill_linear(a, a.numel(), 51, 0, 1.0); ctx.fill_linear(b, b.numel(), 52, 0, 1.0)

        def gemm(ta="N", tb="N"):
            c = ctx.empty(batch * m * n); c.zero_()
            if batch == 1:
                ctx.dgemm(ta, tb, m, n, k, 1.0, a, m if ta == "N" else k, b, k if tb == "N" else n, 0.0, c, m)
            else:
                ctx.dgemm_strided_batched("N", "N", m, n, k, 1.0, a, m, m * k, b, k, k * n, 0.0, c, m, m * n, batch)
            return c
        both(gemm)
        if batch == 1:
            both(lambda: gemm("T", "T"))
        if m == n and batch == 1:
            for uplo in ("U", "L"):
                def syrk():
                    c = ctx.empty(n * n); c.zero_()
                    ctx.dsyrk(uplo, "N", n, k, 1.0, a, n, 0.0, c, n)
                    return c
                both(syrk)
    assert kinds == {0, 1}, "the shape list is meant to cover both the uniform split and stream-K"

    for nb, nx, no in [(264, 720, 21), (600, 200, 60), (45, 77, 7)]:
        sh = ShardedRI(ctx, nb, nx).fill_synthetic()
        cm = ctx.empty(nb * nb); ctx.fill_linear(cm, nb * nb, 3, 0, nb ** -0.5)
        cocc = cm[: nb * no].clone()
        both(lambda: sh.k(cocc, no))
        both(lambda: sh.ao2mo(cm, nb, cm, nb))
        d = sh.dp(cm)
        both(lambda: sh.dp(cm))
        both(lambda: sh.j(d))

    # GEMV column / row chunk partials (workspace slot 1) and the eigen-solver's scratch (slot 0)
    for trans, (m, n) in [("T", (360000, 70)), ("N", (360000, 70)), ("N", (5000, 3)), ("T", (77, 1900))]:
        a = ctx.empty(m * n); ctx.fill_linear(a, m * n, 53, 0, 1.0)
        nx_, ny_ = (n, m) if trans == "N" else (m, n)
        x = ctx.empty(nx_); ctx.fill_linear(x, nx_, 54, 0, 1.0)

        def gemv():
            y = ctx.empty(ny_); y.zero_()
            ctx.dgemv(trans, m, n, 1.0, a, m, x, 1, 0.0, y, 1)
            return y
        both(gemv)
    for n in (3, 65, 300):
        s = ctx.empty(n * n); ctx.fill_linear(s, n * n, 55, 0, 1.0)
        sym = (s.view(n, n) + s.view(n, n).t()).contiguous().view(-1)

        def eig():
            w = ctx.empty(n); z = ctx.empty(n * n)
            ctx.dsyev("V", "U", n, sym.clone(), n, w, z, n)
            return torch.cat([w, z])
        both(eig)
