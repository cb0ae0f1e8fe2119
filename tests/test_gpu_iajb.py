"""GPU parity for the (ia|jb)-type consumers of ri3mo: rb_ri_iajb / rb_host_ri_iajb through the
C ABI vs the CPU oracle (gather + dgemm('T','N') over P on the ri3mo layout of reference src/ri.rs:381-386).
Tolerance: 1e-10 relative, norm-wise and element-wise; symmetric blocks must be bitwise symmetric."""
import numpy as np
import pytest
import torch

from conftest import assert_close_1e10

pytestmark = pytest.mark.gpu


def _dev(ctx, a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(f"cuda:{ctx.device}")


CASES = [
    # np, nl, nr, box A (l0, ll, r0, rl), box B
    (40, 6, 9, (0, 6, 0, 9), (0, 6, 0, 9)),        # whole tensor, symmetric -> SYRK in place
    (41, 6, 9, (0, 6, 0, 9), (0, 6, 0, 9)),        # odd np: pitch repack inside the GEMM core
    (64, 5, 12, (0, 5, 2, 7), (0, 5, 4, 8)),       # two different panels, in place
    (100, 8, 10, (1, 4, 2, 6), (1, 4, 2, 6)),      # partial l range, symmetric -> gather + SYRK
    (100, 8, 10, (0, 3, 0, 10), (3, 5, 1, 4)),     # occ/vir style: different partial boxes
    (333, 16, 24, (2, 9, 3, 17), (0, 16, 5, 11)),  # one gathered, one in place, ragged everything
    (257, 12, 20, (0, 12, 0, 20), (4, 1, 7, 1)),   # single column on one side
    (1, 3, 3, (0, 3, 0, 3), (0, 3, 0, 3)),         # np = 1
]


@pytest.mark.parametrize("case", CASES)
def test_iajb_vs_oracle(ctx, oracle_blas, case):
    np_, nl, nr, ba, bb = case
    mo = oracle_blas.fill_linear(np_ * nl * nr, 31)
    ref = oracle_blas.ri_iajb(np_, mo, nl, ba, mo, nl, bb)
    m, n = ba[1] * ba[3], bb[1] * bb[3]
    mod = _dev(ctx, mo)
    out = ctx.empty(m * n)
    out.fill_(float("nan"))  # beta = 0 must overwrite, never read
    ctx.ri_iajb(np_, mod, np_, nl, nr, ba, mod, np_, nl, nr, bb, 0.0, out, m)
    got = out.cpu().numpy()
    assert_close_1e10(got, ref, f"iajb {case}")
    if ba == bb:
        g = got.reshape((m, m), order="F")
        assert np.array_equal(g, g.T), "symmetric block is not bitwise symmetric"


def test_iajb_beta_ldo_and_padding(ctx, oracle_blas):
    """beta accumulates, ldo > rows leaves the padding rows untouched, ldp > np reads a pitched (sharded) tensor."""
    np_, ldp, nl, nr = 70, 96, 7, 11
    ba, bb = (0, 7, 1, 9), (2, 3, 0, 11)
    m, n, ldo = ba[1] * ba[3], bb[1] * bb[3], ba[1] * ba[3] + 5
    full = oracle_blas.fill_linear(ldp * nl * nr, 32)
    dense = np.ascontiguousarray(full.reshape((ldp, nl, nr), order="F")[:np_].reshape(-1, order="F"))
    ref = oracle_blas.ri_iajb(np_, dense, nl, ba, dense, nl, bb).reshape((m, n), order="F")
    c0 = oracle_blas.fill_linear(ldo * n, 33)
    out = _dev(ctx, c0)
    fd = _dev(ctx, full)
    ctx.ri_iajb(np_, fd, ldp, nl, nr, ba, fd, ldp, nl, nr, bb, -0.5, out, ldo)
    got = out.cpu().numpy().reshape((ldo, n), order="F")
    c0m = c0.reshape((ldo, n), order="F")
    assert_close_1e10(got[:m], ref - 0.5 * c0m[:m], "iajb beta/ldo")
    assert np.array_equal(got[m:], c0m[m:])


def test_iajb_two_tensors_alpha_beta_spin(ctx, oracle_blas):
    """moA != moB (alpha / beta spin blocks with different occupations)."""
    np_ = 90
    nla, nra, nlb, nrb = 5, 14, 4, 15
    a = oracle_blas.fill_linear(np_ * nla * nra, 34)
    b = oracle_blas.fill_linear(np_ * nlb * nrb, 35)
    ba, bb = (0, 5, 0, 14), (1, 3, 2, 13)
    ref = oracle_blas.ri_iajb(np_, a, nla, ba, b, nlb, bb)
    m = ba[1] * ba[3]
    out = ctx.empty(m * bb[1] * bb[3])
    ctx.ri_iajb(np_, _dev(ctx, a), np_, nla, nra, ba, _dev(ctx, b), np_, nlb, nrb, bb, 0.0, out, m)
    assert_close_1e10(out.cpu().numpy(), ref, "iajb two tensors")


def test_iajb_p_shard_additivity(ctx, oracle_blas):
    """P-shard additivity (the multi-GPU contract): the blocks of two row ranges of ri3mo sum to the block of the whole."""
    np_, nl, nr = 200, 6, 10
    ba, bb = (1, 4, 0, 10), (0, 6, 3, 5)
    mo = oracle_blas.fill_linear(np_ * nl * nr, 36)
    ref = oracle_blas.ri_iajb(np_, mo, nl, ba, mo, nl, bb)
    mod = _dev(ctx, mo)
    m, n = ba[1] * ba[3], bb[1] * bb[3]
    out = ctx.empty(m * n)
    split = 87  # odd: the second view is only 8-byte aligned (pitch repack inside the GEMM core)
    # rows [0, split) then rows [split, np) accumulated with beta = 1, each read through the pitched view ldp = np_
    ctx.ri_iajb(split, mod, np_, nl, nr, ba, mod, np_, nl, nr, bb, 0.0, out, m)
    ctx.ri_iajb(np_ - split, mod[split:], np_, nl, nr, ba, mod[split:], np_, nl, nr, bb, 1.0, out, m)
    assert_close_1e10(out.cpu().numpy(), ref, "iajb P-additivity")


def test_iajb_after_ao2mo_occ_vir(ctx, oracle_blas):
    """The real pipeline: occ-vir ao2mo on the device, then (ia|jb) for an (i, j) block pair straight from its output."""
    from rest_tensors_b200.device import ShardedRI
    nb, nx, no = 48, 130, 6
    nv = nb - no
    sh = ShardedRI(ctx, nb, nx).fill_synthetic()
    c = oracle_blas.fill_linear(nb * nb, 3, scale=nb ** -0.5)
    cm = c.reshape((nb, nb), order="F")
    cocc = np.ascontiguousarray(cm[:, :no].reshape(-1, order="F"))
    cvir = np.ascontiguousarray(cm[:, no:].reshape(-1, order="F"))
    mo = sh.ao2mo(_dev(ctx, cocc), no, _dev(ctx, cvir), nv)      # [nx, no, nv]
    ri = oracle_blas.fill_ri3ao_symm(nb, 0, nx)
    mo_ref = oracle_blas.ri_ao2mo_rect(cocc, no, cvir, nv, ri, nb, nx)
    ba = bb = (0, no, 0, nv)
    ref = oracle_blas.ri_iajb(nx, mo_ref, no, ba, mo_ref, no, bb)
    got = sh.iajb(mo, no, nv, ba, bb).cpu().numpy()
    assert_close_1e10(got, ref, "(ia|jb) after occ-vir ao2mo")
    # MP2-like scalar from the block: sum_ijab (ia|jb) [2 (ia|jb) - (ib|ja)]
    g = got.reshape((no, nv, no, nv), order="F"); r = ref.reshape((no, nv, no, nv), order="F")
    e_got = float(np.sum(g * (2.0 * g - g.transpose(0, 3, 2, 1))))
    e_ref = float(np.sum(r * (2.0 * r - r.transpose(0, 3, 2, 1))))
    assert abs(e_got - e_ref) <= 1e-10 * abs(e_ref)


def test_host_iajb_mirror(rt, oracle_blas):
    np_, nl, nr = 75, 6, 9
    mo = oracle_blas.fill_linear(np_ * nl * nr, 37)
    t = rt.RIFull.from_vec([np_, nl, nr], mo)
    for (rla, rra, rlb, rrb) in [((0, 6), (0, 9), (0, 6), (0, 9)), ((1, 4), (2, 8), (0, 6), (3, 5)),
                                 ((0, 6), (1, 3), (0, 6), (6, 9))]:
        ba = (rla[0], rla[1] - rla[0], rra[0], rra[1] - rra[0]); bb = (rlb[0], rlb[1] - rlb[0], rrb[0], rrb[1] - rrb[0])
        ref = oracle_blas.ri_iajb(np_, mo, nl, ba, mo, nl, bb)
        got = t.ri_iajb(rla, rra, rlb, rrb)
        assert got.size == [ba[1] * ba[3], bb[1] * bb[3]]
        assert_close_1e10(got.data, ref, f"host iajb {ba} {bb}")
    with pytest.raises(rt.RestB200Error):
        t.ri_iajb((0, 7), (0, 9), (0, 6), (0, 9))
    # empty boxes are fine and return an empty matrix
    assert t.ri_iajb((2, 2), (0, 9), (0, 6), (0, 9)).size == [0, 54]


def test_iajb_error_codes(ctx):
    from rest_tensors_b200 import RestB200Error
    mo = ctx.empty(10 * 3 * 4)
    out = ctx.empty(144)
    with pytest.raises(RestB200Error):   # box outside
        ctx.ri_iajb(10, mo, 10, 3, 4, (0, 4, 0, 4), mo, 10, 3, 4, (0, 3, 0, 4), 0.0, out, 16)
    with pytest.raises(RestB200Error):   # ldp < np
        ctx.ri_iajb(10, mo, 9, 3, 4, (0, 3, 0, 4), mo, 10, 3, 4, (0, 3, 0, 4), 0.0, out, 12)
    with pytest.raises(RestB200Error):   # ldo < rows
        ctx.ri_iajb(10, mo, 10, 3, 4, (0, 3, 0, 4), mo, 10, 3, 4, (0, 3, 0, 4), 0.0, out, 11)
    # the context is still usable afterwards
    mo.fill_(1.0)
    ctx.ri_iajb(10, mo, 10, 3, 4, (0, 3, 0, 4), mo, 10, 3, 4, (0, 3, 0, 4), 0.0, out, 12)
    assert torch.all(out == 10.0)


# ---------------------------------------------------------------- RPA-type consumer: contraction over the MO pairs ----
PQ_CASES = [
    # np, nl, nr, box (l0, ll, r0, rl), weighted
    (48, 5, 8, (0, 5, 0, 8), False),      # whole tensor in place, SYRK-like
    (48, 5, 8, (0, 5, 0, 8), True),       # weights: scaled copy of the B panel
    (49, 5, 8, (0, 5, 2, 5), True),       # odd np (odd pitch) panel
    (130, 9, 14, (2, 6, 3, 10), False),   # gathered box
    (130, 9, 14, (2, 6, 3, 10), True),
    (300, 4, 50, (1, 2, 0, 50), True),    # np > one tile, ragged
    (7, 3, 3, (1, 1, 2, 1), True),        # a single MO pair: rank-1 update
]


@pytest.mark.parametrize("case", PQ_CASES)
def test_mo_pq_vs_oracle(ctx, oracle_blas, case):
    np_, nl, nr, box, weighted = case
    mo = oracle_blas.fill_linear(np_ * nl * nr, 51)
    w = oracle_blas.fill_linear(box[1] * box[3], 52) if weighted else None
    ref = oracle_blas.ri_mo_pq(mo, np_, mo, np_, nl, box, w)
    mod = _dev(ctx, mo)
    out = ctx.empty(np_ * np_)
    out.fill_(float("nan"))
    ctx.ri_mo_pq(mod, np_, np_, mod, np_, np_, nl, nr, box, None if w is None else _dev(ctx, w), 0.0, out, np_)
    got = out.cpu().numpy()
    assert_close_1e10(got, ref, f"mo_pq {case}")
    g = got.reshape((np_, np_), order="F")
    assert np.array_equal(g, g.T), "moA == moB must give a bitwise symmetric matrix"


def test_mo_pq_row_blocks_and_beta(ctx, oracle_blas):
    """Row blocks of one tensor (what two P-shards exchange): out[P_a, Q_b] with pitched views, beta accumulation,
    ldo > rows; and the four blocks tile the full symmetric matrix."""
    np_, nl, nr = 101, 6, 9
    box = (1, 4, 2, 6)
    mo = oracle_blas.fill_linear(np_ * nl * nr, 53)
    w = oracle_blas.fill_linear(box[1] * box[3], 54)
    full = oracle_blas.ri_mo_pq(mo, np_, mo, np_, nl, box, w).reshape((np_, np_), order="F")
    mod, wd = _dev(ctx, mo), _dev(ctx, w)
    split = 37
    blocks = [(0, split), (split, np_)]
    for (a0, a1) in blocks:
        for (b0, b1) in blocks:
            ma, mb = a1 - a0, b1 - b0
            ldo = ma + 3
            c0 = oracle_blas.fill_linear(ldo * mb, 55)
            out = _dev(ctx, c0)
            ctx.ri_mo_pq(mod[a0:], np_, ma, mod[b0:], np_, mb, nl, nr, box, wd, 2.0, out, ldo)
            got = out.cpu().numpy().reshape((ldo, mb), order="F")
            c0m = c0.reshape((ldo, mb), order="F")
            assert_close_1e10(got[:ma], full[a0:a1, b0:b1] + 2.0 * c0m[:ma], f"mo_pq block {a0}:{a1} x {b0}:{b1}")
            assert np.array_equal(got[ma:], c0m[ma:])


def test_mo_pq_is_iajb_trace_identity(ctx, oracle_blas):
    """Size-independent cross-check of the two consumers: ||G||_F^2 over the (ia|jb) block of a box equals ||Pi||_F^2
    of the unweighted auxiliary-basis matrix of the same box (both are tr(X X^T X X^T))."""
    np_, nl, nr = 96, 7, 12
    box = (0, 7, 0, 12)
    mo = oracle_blas.fill_linear(np_ * nl * nr, 56)
    mod = _dev(ctx, mo)
    m = nl * nr
    g = ctx.empty(m * m); pi = ctx.empty(np_ * np_)
    ctx.ri_iajb(np_, mod, np_, nl, nr, box, mod, np_, nl, nr, box, 0.0, g, m)
    ctx.ri_mo_pq(mod, np_, np_, mod, np_, np_, nl, nr, box, None, 0.0, pi, np_)
    a, b = float(torch.sum(g * g)), float(torch.sum(pi * pi))
    assert abs(a - b) <= 1e-11 * abs(b)


def test_host_mo_pq_mirror(rt, oracle_blas):
    np_, nl, nr = 66, 5, 8
    mo = oracle_blas.fill_linear(np_ * nl * nr, 57)
    t = rt.RIFull.from_vec([np_, nl, nr], mo)
    w = oracle_blas.fill_linear(3 * 5, 58)
    got = t.ri_mo_pq((1, 4), (2, 7), w)
    assert got.size == [np_, np_]
    assert_close_1e10(got.data, oracle_blas.ri_mo_pq(mo, np_, mo, np_, nl, (1, 3, 2, 5), w), "host mo_pq weighted")
    got = t.ri_mo_pq((0, 5), (0, 8))
    assert_close_1e10(got.data, oracle_blas.ri_mo_pq(mo, np_, mo, np_, nl, (0, 5, 0, 8), None), "host mo_pq")
    with pytest.raises(rt.RestB200Error):
        t.ri_mo_pq((0, 6), (0, 8))
    with pytest.raises(rt.RestB200Error):
        t.ri_mo_pq((0, 5), (0, 8), np.ones(3))


def test_symmetric_case_with_beta_keeps_blas_semantics(ctx, oracle_blas):
    """moA == moB with identical boxes and beta != 0: out need not be symmetric on entry, so the result must be
    beta*out + G element by element (no triangle is rebuilt from the other one)."""
    np_, nl, nr = 60, 4, 6
    box = (0, 4, 1, 5)
    m = box[1] * box[3]
    mo = oracle_blas.fill_linear(np_ * nl * nr, 61)
    g = oracle_blas.ri_iajb(np_, mo, nl, box, mo, nl, box)
    c0 = oracle_blas.fill_linear(m * m, 62)       # not symmetric
    out = _dev(ctx, c0)
    mod = _dev(ctx, mo)
    ctx.ri_iajb(np_, mod, np_, nl, nr, box, mod, np_, nl, nr, box, 1.0, out, m)
    assert_close_1e10(out.cpu().numpy(), g + c0, "iajb symmetric box, beta = 1")


def test_consumers_at_config_c_scale(ctx, oracle_blas):
    """Bench-sized inputs (config C: naux = 1700, nocc = 60, nvir = 540) straight from the device-side occ-vir ao2mo:
    an (ia|jb) block of 12 x 12 occupied orbitals (6480 x 6480, K = 1700) in full against the oracle + OpenBLAS, and the
    size-independent cross-check ||G||_F^2 = ||Pi||_F^2 between the two consumers on a whole-l box."""
    from rest_tensors_b200.device import ShardedRI
    nb, nx, no = 600, 1700, 60
    nv = nb - no
    sh = ShardedRI(ctx, nb, nx).fill_synthetic()
    c = ctx.empty(nb * nb); ctx.fill_linear(c, nb * nb, 3, 0, nb ** -0.5)
    mo = sh.ao2mo(c[: nb * no], no, c[nb * no:], nv)              # [1700, 60, 540] in HBM
    del sh
    mo_h = mo.cpu().numpy()
    li = 12
    ba, bb = (0, li, 0, nv), (li, li, 0, nv)
    m = li * nv
    g = ctx.empty(m * m)
    ctx.ri_iajb(nx, mo, nx, no, nv, ba, mo, nx, no, nv, bb, 0.0, g, m)
    assert_close_1e10(g.cpu().numpy(), oracle_blas.ri_iajb(nx, mo_h, no, ba, mo_h, no, bb), "(ia|jb) off-diagonal block, config C")
    ctx.ri_iajb(nx, mo, nx, no, nv, ba, mo, nx, no, nv, ba, 0.0, g, m)
    gd = g.cpu().numpy()
    assert_close_1e10(gd, oracle_blas.ri_iajb(nx, mo_h, no, ba, mo_h, no, ba), "(ia|jb) diagonal block, config C")
    gm = gd.reshape((m, m), order="F")
    assert np.array_equal(gm, gm.T)
    # whole-l box of 4 virtuals: both consumers are functions of X = mo[:, :, 0:4] (1700 x 240)
    box = (0, no, 0, 4)
    k = no * 4
    g2 = ctx.empty(k * k); pi = ctx.empty(nx * nx)
    ctx.ri_iajb(nx, mo, nx, no, nv, box, mo, nx, no, nv, box, 0.0, g2, k)
    ctx.ri_mo_pq(mo, nx, nx, mo, nx, nx, no, nv, box, None, 0.0, pi, nx)
    a, b = float(torch.sum(g2 * g2)), float(torch.sum(pi * pi))
    assert abs(a - b) <= 1e-11 * abs(b)
    assert_close_1e10(pi.cpu().numpy(), oracle_blas.ri_mo_pq(mo_h, nx, mo_h, nx, no, box, None), "Pi, config C rows")
