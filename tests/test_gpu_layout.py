"""GPU parity, bit-exact half: pack/unpack, per-slab symmetric pack, sub-box copies, transposes, axpy family and the
synthetic generators -- CUDA (through the C ABI) vs the CPU oracle on the same inputs.  Everything here must match
BIT FOR BIT (np.array_equal)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _dev(ctx, a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(f"cuda:{ctx.device}")


# ---------------------------------------------------------------- golden vectors through the product ----
def test_golden_pack_unpack_transpose(rt, golden):
    g = golden["GV4"]
    assert rt.MatrixFull.from_vec([4, 4], g["full"]).to_matrixupper().data.tolist() == g["packed"]
    for case in golden["GV5"]["cases"]:
        assert rt.MatrixUpper.from_vec(6, case["packed"]).to_matrixfull().data.tolist() == case["full_colmajor"]
    g = golden["GV7"]
    t = rt.MatrixFull.from_vec(g["size"], g["data"]).transpose()
    assert t.size == [4, 3]
    assert t.data.reshape((4, 3), order="F")[:, 2].tolist() == g["transposed_column_2"]


def test_golden_ri_transposes(rt, golden):
    g = golden["GV10"]
    r = rt.RIFull.from_vec(g["size"], g["data"])
    assert r.transpose_jik().data.tolist() == g["jik"] and r.transpose_jik().size == [2, 3, 2]
    assert r.transpose_jki().data.tolist() == g["jki"] and r.transpose_jki().size == [2, 2, 3]
    assert r.transpose_kji().data.tolist() == g["kji"] and r.transpose_kji().size == [2, 2, 3]
    assert r.transpose_ikj().data.tolist() == g["ikj"] and r.transpose_ikj().size == [3, 2, 2]


# ---------------------------------------------------------------- pack / unpack ----
@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 100, 264, 511, 600, 1000])
def test_pack_unpack_vs_oracle(rt, oracle, n):
    full = oracle.fill_linear(n * n, 4)
    packed_ref = oracle.to_matrixupper(full, n)
    packed = rt.MatrixFull.from_vec([n, n], full).to_matrixupper()
    assert packed.size == n * (n + 1) // 2
    assert np.array_equal(packed.data, packed_ref)
    unpacked = packed.to_matrixfull()
    assert unpacked.size == [n, n]
    assert np.array_equal(unpacked.data, oracle.to_matrixfull(packed_ref))
    # round trip: pack(unpack(p)) == p ; unpack is symmetric
    assert np.array_equal(unpacked.to_matrixupper().data, packed_ref)
    m = unpacked.data.reshape((n, n), order="F")
    assert np.array_equal(m, m.T)


@pytest.mark.parametrize("n", [4000, 8000])
def test_pack_unpack_full_size_properties(ctx, n):
    """BASELINE config E sizes on device buffers: round trip, symmetry and a checksum of checksums."""
    np_ = n * (n + 1) // 2
    packed = ctx.empty(np_)
    ctx.fill_linear(packed, np_, 4, 0, 1.0)
    full = ctx.empty(n * n)
    ctx.unpack_upper(packed, n, full)
    back = ctx.empty(np_)
    ctx.pack_upper(full, n, back)
    assert torch.equal(back, packed)
    f2 = full.view(n, n)  # column-major n x n viewed row-major == transpose; symmetric either way
    assert torch.equal(f2, f2.t())
    # diagonal of the full matrix == packed[j(j+1)/2 + j]
    j = torch.arange(n, device=packed.device)
    assert torch.equal(torch.diagonal(f2), packed[j * (j + 1) // 2 + j])


def test_ri_pack_symm(rt, oracle):
    for nao, naux in [(7, 5), (33, 9), (100, 40)]:
        ri = oracle.fill_linear(nao * nao * naux, 2)
        out = rt.RIFull.from_vec([nao, nao, naux], ri).rifull_to_matfull_symm()
        assert out.size == [nao * (nao + 1) // 2, naux]
        assert np.array_equal(out.data, oracle.rifull_to_matfull_symm(ri, nao, naux))
    big = rt.RIFull.new([600, 600, 3], 0.0)
    big.data[:] = oracle.fill_linear(big.data.size, 2)
    assert np.array_equal(big.rifull_to_matfull_symm().data, oracle.rifull_to_matfull_symm(big.data, 600, 3))


# ---------------------------------------------------------------- transposes ----
@pytest.mark.parametrize("shape", [(1, 1, 1), (3, 2, 2), (9, 3, 2), (33, 65, 7), (100, 100, 17), (64, 1, 130)])
def test_ri_transposes_vs_oracle(rt, oracle, shape):
    i, j, k = shape
    d = oracle.fill_linear(i * j * k, 5)
    r = rt.RIFull.from_vec([i, j, k], d)
    for which, fn in enumerate([r.transpose_jik, r.transpose_jki, r.transpose_kji, r.transpose_ikj]):
        assert np.array_equal(fn().data, oracle.ri_transpose(d, i, j, k, which)), (shape, which)
    # involutions: jik(jik(x)) == x, kji(kji(x)) == x, ikj(ikj(x)) == x
    assert np.array_equal(r.transpose_jik().transpose_jik().data, d)
    assert np.array_equal(r.transpose_kji().transpose_kji().data, d)
    assert np.array_equal(r.transpose_ikj().transpose_ikj().data, d)


@pytest.mark.parametrize("shape", [(3, 4), (1, 7), (129, 65), (600, 264)])
def test_matrix_transpose_vs_oracle(rt, oracle, shape):
    r, c = shape
    d = oracle.fill_linear(r * c, 6)
    t = rt.MatrixFull.from_vec([r, c], d).transpose()
    assert np.array_equal(t.data, oracle.matrix_transpose(d, r, c))
    assert np.array_equal(t.transpose().data, d)


# ---------------------------------------------------------------- sub-box copies ----
def test_copy_mm(rt, oracle):
    src = oracle.fill_linear(13 * 11, 7)
    for (xl, yl, fxs, fys, txs, tys) in [(5, 4, 2, 3, 1, 0), (13, 11, 0, 0, 0, 0), (1, 1, 12, 10, 16, 8), (0, 3, 0, 0, 0, 0),
                                         (6, 2, 4, 4, 2, 2)]:
        dst_ref = oracle.fill_linear(17 * 9, 8)
        dst = rt.MatrixFull.from_vec([17, 9], dst_ref.copy())
        if yl + tys > 9:
            continue
        oracle.copy_mm(xl, yl, src, 13, 11, fxs, fys, dst_ref, 17, 9, txs, tys)
        dst.copy_from_matr((txs, txs + xl), (tys, tys + yl), rt.MatrixFull.from_vec([13, 11], src), (fxs, fxs + xl),
                           (fys, fys + yl))
        assert np.array_equal(dst.data, dst_ref)


@pytest.mark.parametrize("mod", [0, 1, 2])
def test_copy_mr_and_rm(rt, oracle, mod):
    X, Y, Z = 6, 5, 4
    dims = {0: (X, Y, Z), 1: (X, Z, Y), 2: (Y, Z, X)}[mod]   # extents of (x1, x2, x3) for this mode
    src = oracle.fill_linear(9 * 8, 9)
    for (l1, l2, f1, f2, t1, t2, x3) in [(3, 2, 1, 2, 1, 1, 2), (dims[0], dims[1], 0, 0, 0, 0, 0), (1, 1, 8, 7, 0, 0, dims[2] - 1)]:
        if l1 + t1 > dims[0] or l2 + t2 > dims[1] or x3 >= dims[2] or l1 + f1 > 9 or l2 + f2 > 8:
            continue
        t_ref = oracle.fill_linear(X * Y * Z, 10)
        t = rt.RIFull.from_vec([X, Y, Z], t_ref.copy())
        oracle.copy_mr(l1, l2, src, 9, 8, f1, f2, t_ref, X, Y, Z, t1, t2, x3, mod)
        t.copy_from_matr((t1, t1 + l1), (t2, t2 + l2), x3, mod, rt.MatrixFull.from_vec([9, 8], src), (f1, f1 + l1),
                         (f2, f2 + l2))
        assert np.array_equal(t.data, t_ref), (mod, l1, l2)
        # and back out of the tensor (copy_rm, reachable through external_libs::matr_copy_from_ri)
        m_ref = oracle.fill_linear(9 * 8, 11)
        m = m_ref.copy()
        oracle.copy_rm(l1, l2, t_ref, X, Y, Z, t1, t2, x3, mod, m_ref, 9, 8, f1, f2)
        rt.matr_copy_from_ri(t.data, t.size, (t1, t1 + l1), (t2, t2 + l2), x3, mod, m, [9, 8], (f1, f1 + l1), (f2, f2 + l2))
        assert np.array_equal(m, m_ref), (mod, l1, l2)
    # unknown mod is a no-op (restmatr.f90:227-237)
    t0 = oracle.fill_linear(X * Y * Z, 10)
    t = rt.RIFull.from_vec([X, Y, Z], t0.copy())
    t.copy_from_matr((0, 2), (0, 2), 0, 7, rt.MatrixFull.from_vec([9, 8], src), (0, 2), (0, 2))
    assert np.array_equal(t.data, t0)


def test_copy_rr(rt, oracle):
    f = oracle.fill_linear(7 * 6 * 5, 12)
    for (xl, yl, zl, fs, ts) in [((3, 2, 2), None, None, (1, 2, 1), (0, 1, 3)), ((7, 6, 5), None, None, (0, 0, 0), (0, 0, 0)),
                                 ((2, 6, 1), None, None, (5, 0, 4), (6, 0, 0))]:
        xl, yl, zl = xl
        t_ref = oracle.fill_linear(8 * 7 * 6, 13)
        t = rt.RIFull.from_vec([8, 7, 6], t_ref.copy())
        oracle.copy_rr(xl, yl, zl, f, 7, 6, 5, fs[0], fs[1], fs[2], t_ref, 8, 7, 6, ts[0], ts[1], ts[2])
        t.copy_from_ri((ts[0], ts[0] + xl), (ts[1], ts[1] + yl), (ts[2], ts[2] + zl), rt.RIFull.from_vec([7, 6, 5], f),
                       (fs[0], fs[0] + xl), (fs[1], fs[1] + yl), (fs[2], fs[2] + zl))
        assert np.array_equal(t.data, t_ref)


def test_device_copies_large(ctx, oracle):
    """device-pointer copies on a box large enough to exercise the vectorised and the scalar kernels"""
    fx, fy, fz, tx, ty, tz = 130, 70, 9, 140, 80, 11
    f = oracle.fill_linear(fx * fy * fz, 14)
    for (xl, yl, zl, fs, ts) in [(128, 64, 8, (2, 4, 1), (4, 8, 2)), (127, 63, 7, (1, 3, 0), (3, 5, 1))]:
        t_ref = oracle.fill_linear(tx * ty * tz, 15)
        td = _dev(ctx, t_ref)
        oracle.copy_rr(xl, yl, zl, f, fx, fy, fz, fs[0], fs[1], fs[2], t_ref, tx, ty, tz, ts[0], ts[1], ts[2])
        ctx.copy_rr(xl, yl, zl, _dev(ctx, f), fx, fy, fz, fs[0], fs[1], fs[2], td, tx, ty, tz, ts[0], ts[1], ts[2])
        assert np.array_equal(td.cpu().numpy(), t_ref)
    with pytest.raises(Exception):
        ctx.copy_rr(10, 10, 10, _dev(ctx, f), fx, fy, fz, 125, 0, 0, _dev(ctx, f), fx, fy, fz, 0, 0, 0)  # box outside


# ---------------------------------------------------------------- axpy family ----
def test_axpy_family_bit_exact(rt, oracle):
    n = 12345
    for op, a, b in [(0, 0.0, 0.37), (1, -1.25, 0.3333333333333333), (2, 1.0 / 3.0, 0.0), (3, 0.0, 0.0), (4, 0.0, 0.0)]:
        c = oracle.fill_linear(n, 16); p = oracle.fill_linear(n, 17)
        c_ref = c.copy()
        oracle.axpy(op, c_ref, p, a, b)
        m = rt.MatrixFull.from_vec([n, 1], c.copy()); q = rt.MatrixFull.from_vec([n, 1], p)
        [lambda: m.self_scaled_add(q, b), lambda: m.self_general_add(q, a, b), lambda: m.self_multiple(a),
         lambda: m.self_add(q), lambda: m.self_sub(q)][op]()
        assert np.array_equal(m.data, c_ref), op
    r = rt.RIFull.from_vec([5, 4, 3], oracle.fill_linear(60, 18))
    ref = r.data.copy()
    oracle.axpy(0, ref, oracle.fill_linear(60, 19), 0.0, -2.7)
    r.self_scaled_add(rt.RIFull.from_vec([5, 4, 3], oracle.fill_linear(60, 19)), -2.7)
    assert np.array_equal(r.data, ref)
    # MatrixUpper +/- truncates to the shorter operand (matrixupper.rs:395-420)
    u = rt.MatrixUpper.from_vec(6, np.arange(6.0)) + rt.MatrixUpper.from_vec(3, np.ones(3))
    assert u.data.tolist() == [1.0, 2.0, 3.0, 3.0, 4.0, 5.0]
    u = rt.MatrixUpper.from_vec(6, np.arange(6.0)) - rt.MatrixUpper.from_vec(6, np.ones(6))
    assert u.data.tolist() == [-1.0, 0.0, 1.0, 2.0, 3.0, 4.0]
    out = rt.MatrixFull.from_vec([2, 2], [1.0, 2.0, 3.0, 4.0]).scaled_add(rt.MatrixFull.from_vec([2, 2], [1.0, 1.0, 1.0, 1.0]), 0.5)
    assert out.data.tolist() == [1.5, 2.5, 3.5, 4.5]


# ---------------------------------------------------------------- synthetic generators ----
def test_device_generators_match_oracle(ctx, oracle):
    v = ctx.empty(10007)
    ctx.fill_linear(v, 10007, 3, 5, 0.1)
    assert np.array_equal(v.cpu().numpy(), oracle.fill_linear(10007, 3, 5, 0.1))
    nb, p_lo, p_hi = 37, 3, 9
    a = ctx.empty(nb * nb * (p_hi - p_lo))
    ctx.fill_ri3ao_symm(a, nb, p_lo, p_hi, 1, 1.0)
    assert np.array_equal(a.cpu().numpy(), oracle.fill_ri3ao_symm(nb, p_lo, p_hi, 1, 1.0))


def test_einsum_helpers(rt, oracle):
    """"ij,j->ij" and "i,j->ij" are single multiplies (bit-exact), "ip,ip->p" a column dot (1e-10)."""
    from conftest import assert_close_1e10
    for ni, nj in [(1, 1), (7, 5), (1000, 33), (4097, 3), (64, 3000)]:
        a = oracle.fill_linear(ni * nj, 51); b = oracle.fill_linear(ni * nj, 52)
        vj = oracle.fill_linear(nj, 53); vi = oracle.fill_linear(ni, 54)
        A = rt.MatrixFull.from_vec([ni, nj], a); B = rt.MatrixFull.from_vec([ni, nj], b)
        for fn in (rt._einsum_01_rayon, rt._einsum_01_serial):
            assert np.array_equal(fn(A.to_matrixfullslice(), vj).data, oracle.einsum_01(a, vj, ni, nj))
        for fn in (rt._einsum_02_rayon, rt._einsum_02_serial):
            assert_close_1e10(fn(A.to_matrixfullslice(), B.to_matrixfullslice()), oracle.einsum_02(a, b, ni, nj), "ip,ip->p")
        for fn in (rt._einsum_03, rt._einsum_03_forvec):
            assert np.array_equal(fn(vi, vj).data, oracle.einsum_03(vi, vj, ni, nj))
    # fewer scale factors than columns: only the leading columns are produced (zip semantics)
    A = rt.MatrixFull.from_vec([5, 4], oracle.fill_linear(20, 55))
    out = rt._einsum_01_rayon(A, np.array([2.0, -1.0]))
    assert out.size == [5, 2] and np.array_equal(out.data, np.concatenate([A.data[:5] * 2.0, A.data[5:10] * -1.0]))
    # dispatcher
    g = rt._einsum_general(A, rt.MatrixFull.from_vec([4, 1], np.arange(4.0)), "ij,j->ij")
    assert g.size == [5, 4] and np.array_equal(g.data, oracle.einsum_01(A.data, np.arange(4.0), 5, 4))
    assert rt._einsum_general(A, A, "ip,ip->p").size == [4, 1]
    with pytest.raises(ValueError):
        rt._einsum_general(A, A, "ijk->i")
    assert rt._einsum_03(np.zeros(0), np.ones(3)).size == [0, 3]


def test_c_abi_error_codes_and_recovery(ctx, rt):
    """Device API: invalid arguments return RB_ERR_INVALID with a message (never a CUDA error / abort), and the
    context keeps working afterwards.  The Rust wrappers panic on the same conditions before the FFI call."""
    import ctypes as C
    from rest_tensors_b200._lib import lib, last_error, RestB200Error
    a = ctx.empty(64); b = ctx.empty(64); c = ctx.empty(64)
    P = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    bad = [
        (lambda: lib.rb_dgemm(ctx.h, b"N", b"N", 8, 8, 8, 1.0, P(a), 4, P(b), 8, 0.0, P(c), 8), "lda"),       # lda < m
        (lambda: lib.rb_dgemm(ctx.h, b"N", b"N", 8, 8, 8, 1.0, P(a), 8, P(b), 8, 0.0, P(c), 4), "ldc"),       # ldc < m
        (lambda: lib.rb_dgemm(ctx.h, b"Q", b"N", 8, 8, 8, 1.0, P(a), 8, P(b), 8, 0.0, P(c), 8), "trans"),
        (lambda: lib.rb_dsyrk(ctx.h, b"X", b"N", 8, 8, 1.0, P(a), 8, 0.0, P(c), 8), "uplo"),
        (lambda: lib.rb_ri_ao2mo(ctx.h, P(a), 2, P(a), 2, P(b), P(c), 2, 4, 3), "out_ldp"),                   # ldp < nx
        (lambda: lib.rb_ri_transpose(ctx.h, P(a), 2, 2, 2, 9, P(c)), "which"),
        (lambda: lib.rb_copy_mm(ctx.h, 5, 5, P(a), 8, 8, 6, 0, P(c), 8, 8, 0, 0), "outside"),
        (lambda: lib.rb_unpack_upper(ctx.h, None, 4, P(c)), "NULL"),
    ]
    for call, word in bad:
        st = call()
        assert st == 1, f"expected RB_ERR_INVALID, got {st}"
        assert word.lower() in last_error().lower(), (word, last_error())
    with pytest.raises(RestB200Error):
        ctx.dgemm("N", "N", 8, 8, 8, 1.0, a, 4, b, 8, 0.0, c, 8)
    # still healthy
    a.fill_(1.0); b.fill_(2.0)
    ctx.dgemm("N", "N", 8, 8, 8, 1.0, a, 8, b, 8, 0.0, c, 8)
    assert torch.all(c == 16.0)
    node = C.c_int(-5)
    assert lib.rb_bind_host_to_device_numa(0, C.byref(node)) == 0 and node.value >= -1


def test_host_wrappers_large_pageable_operands(rt, oracle, oracle_blas):
    """Operands >= 16 MB in pageable memory (numpy arrays) take the pipelined pinned-bounce path of the host wrappers
    (dense and pitched, several 64 MB pieces); results must be what the small direct path gives: bit-exact layout ops,
    1e-10 GEMM with leading dimensions larger than the rows."""
    from conftest import assert_close_1e10
    n = 4300                                         # packed 9.2 M doubles (2 pieces), full 18.5 M (3 pieces)
    packed = oracle.fill_linear(n * (n + 1) // 2, 61)
    full = rt.MatrixUpper.from_vec(packed.size, packed).to_matrixfull()
    assert np.array_equal(full.data, oracle.to_matrixfull(packed))
    assert np.array_equal(full.to_matrixupper().data, packed)
    r, c = 3000, 2500
    a = oracle.fill_linear(r * c, 62)
    assert np.array_equal(rt.MatrixFull.from_vec([r, c], a).transpose().data, oracle.matrix_transpose(a, r, c))
    # sub-block GEMM through general_dgemm_f_: pitched uploads / download of blocks inside larger matrices
    ra, ca, rb, cb, rc, cc = 2600, 1700, 1700, 2300, 2700, 2400
    m, k, nn = 2500, 1600, 2200
    A = rt.MatrixFull.from_vec([ra, ca], oracle.fill_linear(ra * ca, 63))
    B = rt.MatrixFull.from_vec([rb, cb], oracle.fill_linear(rb * cb, 64))
    Cm = rt.MatrixFull.from_vec([rc, cc], oracle.fill_linear(rc * cc, 65))
    c_ref = Cm.data.copy()
    oracle_blas.general_dgemm_f(A.data, [ra, ca], (50, 50 + m), (60, 60 + k), "N", B.data, [rb, cb], (70, 70 + k), (30, 30 + nn), "N",
                                c_ref, [rc, cc], (100, 100 + m), (90, 90 + nn), 0.7, 0.2)
    rt._dgemm(A, ((50, 50 + m), (60, 60 + k)), "N", B, ((70, 70 + k), (30, 30 + nn)), "N", Cm, ((100, 100 + m), (90, 90 + nn)), 0.7, 0.2)
    assert_close_1e10(Cm.data, c_ref, "large sub-block dgemm, pageable operands")


def test_host_entry_points_from_many_threads(rt, oracle_blas):
    """REST calls the wrappers from rayon worker threads (per-slab work inside par_iter_auxbas): the host-pointer entry
    points must be safe to call concurrently (they serialise on the default context) and keep every result intact."""
    import threading
    from conftest import assert_close_1e10
    m, k, n = 150, 90, 70
    jobs = []
    for t in range(8):
        a = oracle_blas.fill_linear(m * k, 70 + t); b = oracle_blas.fill_linear(k * n, 90 + t)
        ref = np.zeros(m * n); oracle_blas.dgemm("N", "N", m, n, k, 1.0, a, m, b, k, 0.0, ref, m)
        jobs.append((a, b, ref))
    errors = []

    def work(t):
        try:
            a, b, ref = jobs[t]
            for _ in range(5):
                c = rt._dgemm_full_new(rt.MatrixFull.from_vec([m, k], a), "N", rt.MatrixFull.from_vec([k, n], b), "N", 1.0, 0.0)
                assert_close_1e10(c.data, ref, f"thread {t}")
                p = rt.MatrixFull.from_vec([n, n], ref[: n * n].copy()).to_matrixupper()
                assert p.data.size == n * (n + 1) // 2
        except Exception as exc:  # noqa: BLE001
            errors.append(f"thread {t}: {exc}")

    threads = [threading.Thread(target=work, args=(t,)) for t in range(8)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors


# ---------------------------------------------------------------- bulk-tensor (TMA) forms, device buffers ----
def _ri_transposed_view(x, i, j, k, which):
    """torch view of what transpose `which` must produce (ri.rs:227-294), from the column-major [i, j, k] buffer x"""
    t = x.view(k, j, i)                       # element (i, j, k) of the column-major tensor == t[k][j][i]
    perm = [(0, 2, 1), (2, 0, 1), (2, 1, 0), (1, 0, 2)][which]   # jik, jki, kji, ikj (slowest output index first)
    return t.permute(*perm).contiguous().view(-1)


@pytest.mark.parametrize("shape", [(64, 64, 16), (256, 130, 6), (130, 258, 10), (66, 1000, 4), (1000, 66, 6), (600, 600, 8),
                                   (2, 40000, 2), (40000, 2, 2)])
@pytest.mark.parametrize("path", [0, 1])
def test_tma_ri_transposes_bit_exact(ctx, shape, path):
    """Even extents and aligned buffers: the 32-byte (LDG/STG.256) kernels (path 0, default) and the bulk-tensor kernels
    (path 1: TMA load -> shared memory -> TMA store): ragged tiles at every edge, thin tensors, all four permutations,
    compared bit for bit with torch's permute."""
    i, j, k = shape
    n = i * j * k
    x = ctx.empty(n); ctx.fill_linear(x, n, 21, 0, 1.0)
    u = ctx.empty(n)
    ctx.set_layout_path(path)
    try:
        before = ctx.tma_layout_launches
        for which in range(4):
            u.fill_(float("nan"))
            ctx.ri_transpose(x, i, j, k, which, u)
            assert torch.equal(u, _ri_transposed_view(x, i, j, k, which)), (shape, which, path)
        assert ctx.tma_layout_launches == before + (4 if path else 0), "path 1 must run aligned transposes on the bulk-tensor kernels"
    finally:
        ctx.set_layout_path(0)


@pytest.mark.parametrize("path", [0, 1])
@pytest.mark.parametrize("rows,cols", [(4000, 2000), (1002, 514), (64, 8192), (8192, 66), (600, 264), (1004, 516), (68, 4100)])
def test_tma_matrix_transpose_bit_exact(ctx, rows, cols, path):
    x = ctx.empty(rows * cols); ctx.fill_linear(x, rows * cols, 22, 0, 1.0)
    u = ctx.empty(rows * cols); u.fill_(float("nan"))
    ctx.set_layout_path(path)
    try:
        before = ctx.tma_layout_launches
        ctx.matrix_transpose(x, rows, cols, u)
        assert torch.equal(u.view(rows, cols), x.view(cols, rows).t())
        assert ctx.tma_layout_launches == before + (1 if path else 0)
    finally:
        ctx.set_layout_path(0)


@pytest.mark.parametrize("path", [0, 1])
def test_tma_sub_box_copies_bit_exact(ctx, path):
    """copy_rr / copy_mm with even starts (16-byte aligned box corners): bulk-tensor copy; everything outside the box
    must stay untouched.  Odd starts or an odd unit-stride extent (a bulk-tensor store clips that extent in 16-byte units:
    it would overwrite the element just past an odd box) fall back to the plain kernel and must agree too."""
    fx, fy, fz, tx, ty, tz = 600, 520, 12, 640, 530, 14
    f = ctx.empty(fx * fy * fz); ctx.fill_linear(f, f.numel(), 23, 0, 1.0)
    for (xl, yl, zl, fs, ts, tma) in [(500, 500, 10, (50, 10, 1), (20, 4, 2), True), (600, 520, 12, (0, 0, 0), (0, 0, 0), True),
                                      (2, 500, 12, (598, 3, 0), (0, 7, 1), False), (300, 1, 12, (0, 519, 0), (340, 0, 2), False),
                                      (501, 500, 10, (50, 10, 1), (20, 4, 2), False), (500, 500, 10, (51, 10, 1), (20, 4, 2), False),
                                      (500, 499, 9, (50, 11, 1), (20, 5, 2), True)]:
        t = ctx.empty(tx * ty * tz); ctx.fill_linear(t, t.numel(), 24, 0, 1.0)
        ref = t.clone()
        ref.view(tz, ty, tx)[ts[2]:ts[2] + zl, ts[1]:ts[1] + yl, ts[0]:ts[0] + xl] = \
            f.view(fz, fy, fx)[fs[2]:fs[2] + zl, fs[1]:fs[1] + yl, fs[0]:fs[0] + xl]
        ctx.set_layout_path(path)
        try:
            before = ctx.tma_layout_launches
            ctx.copy_rr(xl, yl, zl, f, fx, fy, fz, fs[0], fs[1], fs[2], t, tx, ty, tz, ts[0], ts[1], ts[2])
            assert torch.equal(t, ref), (xl, yl, zl, fs, ts, path)
            assert (ctx.tma_layout_launches == before + 1) == (tma and path == 1), (xl, yl, zl, fs, ts)
        finally:
            ctx.set_layout_path(0)
    # boxes whose corners and extents are multiples of 4 take the 32-byte kernels (path 0); whole slabs collapse into one run
    for (xl, yl, zl, fs, ts) in [(496, 500, 10, (52, 10, 1), (20, 4, 2)), (600, 520, 3, (0, 0, 2), (0, 0, 0))]:
        if (xl, yl) == (600, 520):
            t = ctx.empty(fx * fy * 5); ctx.fill_linear(t, t.numel(), 24, 0, 1.0)
            ref = t.clone(); ref.view(5, fy, fx)[0:3] = f.view(fz, fy, fx)[2:5]
            ctx.copy_rr(xl, yl, zl, f, fx, fy, fz, 0, 0, 2, t, fx, fy, 5, 0, 0, 0)
        else:
            t = ctx.empty(tx * ty * tz); ctx.fill_linear(t, t.numel(), 24, 0, 1.0)
            ref = t.clone()
            ref.view(tz, ty, tx)[ts[2]:ts[2] + zl, ts[1]:ts[1] + yl, ts[0]:ts[0] + xl] = \
                f.view(fz, fy, fx)[fs[2]:fs[2] + zl, fs[1]:fs[1] + yl, fs[0]:fs[0] + xl]
            ctx.copy_rr(xl, yl, zl, f, fx, fy, fz, fs[0], fs[1], fs[2], t, tx, ty, tz, ts[0], ts[1], ts[2])
        assert torch.equal(t, ref), (xl, yl, zl)
    n = 4000
    a = ctx.empty(n * n); ctx.fill_linear(a, n * n, 25, 0, 1.0)
    b = ctx.empty(n * n); b.zero_()
    ctx.copy_mm(n - 100, n - 200, a, n, n, 100, 200, b, n, n, 0, 0)
    ref = torch.zeros_like(b)
    ref.view(n, n)[0:n - 200, 0:n - 100] = a.view(n, n)[200:n, 100:n]
    assert torch.equal(b, ref)


@pytest.mark.parametrize("n", [64, 68, 128, 252, 600, 1000, 1028, 4000])
def test_pack_unpack_256bit_kernels_bit_exact(ctx, n):
    """n % 4 == 0 with 32-byte aligned buffers runs the LDG/STG.256 pack / unpack kernels: compare with an independent torch
    construction of the packed <-> full maps (matrixupper.rs:330-373, matrixfull.rs:638-646), bit for bit, including
    the mirrored triangle, ragged edge tiles (n % 64 != 0) and the 4 x 4 blocks on the diagonal."""
    np_ = n * (n + 1) // 2
    packed = ctx.empty(np_); ctx.fill_linear(packed, np_, 4, 0, 1.0)
    full = ctx.empty(n * n); full.fill_(float("nan"))
    ctx.unpack_upper(packed, n, full)
    jj, ii = torch.meshgrid(torch.arange(n, device=packed.device), torch.arange(n, device=packed.device), indexing="ij")  # [col j][row i]
    lo, hi = torch.minimum(ii, jj), torch.maximum(ii, jj)
    want = packed[hi * (hi + 1) // 2 + lo]                   # full[i + j n] viewed as [j][i]
    assert torch.equal(full.view(n, n), want)
    back = ctx.empty(np_); back.fill_(float("nan"))
    ctx.pack_upper(full, n, back)
    assert torch.equal(back, packed)
    # pack must read the upper triangle only: poison the strictly lower part
    poisoned = full.clone().view(n, n)
    poisoned[ii > jj] = float("nan")
    ctx.pack_upper(poisoned.reshape(-1), n, back)
    assert torch.equal(back, packed)


def test_ri_pack_symm_256bit_bit_exact(ctx):
    nao, naux = 128, 7
    ri = ctx.empty(nao * nao * naux); ctx.fill_linear(ri, ri.numel(), 2, 0, 1.0)
    out = ctx.empty(nao * (nao + 1) // 2 * naux); out.fill_(float("nan"))
    ctx.ri_pack_symm(ri, nao, naux, out)
    jj, ii = torch.meshgrid(torch.arange(nao, device=ri.device), torch.arange(nao, device=ri.device), indexing="ij")
    sel = (ii <= jj)
    for p in range(naux):
        slab = ri.view(naux, nao, nao)[p]                    # [j][i]
        assert torch.equal(out.view(naux, -1)[p], slab[sel])  # row-major walk of [j][i] with i <= j == packed order j(j+1)/2 + i


def test_axpy_256bit_body_and_tail_bit_exact(ctx):
    for n in (4096, 4099, 100003):
        c = ctx.empty(n); p = ctx.empty(n)
        ctx.fill_linear(c, n, 16, 0, 1.0); ctx.fill_linear(p, n, 17, 0, 1.0)
        want = c + p * 0.37                                   # torch: separate multiply and add kernels -> unfused, like the reference
        ctx.self_scaled_add(c, p, 0.37, n)
        assert torch.equal(c, want), n
        want = c * (-1.25) + p * (1.0 / 3.0)
        ctx.self_general_add(c, p, -1.25, 1.0 / 3.0, n)
        assert torch.equal(c, want), n
