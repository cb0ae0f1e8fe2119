"""CPU-side checks: the C-ABI library loads and exports every symbol include/rest_b200.h declares, the product
fails loudly without a GPU (no CPU fallback), and the host-side mirror keeps the reference's metadata / panics."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "rest_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(?:int|void|int64_t|const char \*)\s*\*?\s*([a-z_0-9]+)\s*\(", src)
    return sorted(set(n for n in names if n.startswith("rb_") or n.endswith("_")))


def test_library_exports_every_declared_symbol(rt):
    names = _header_functions()
    assert len(names) > 60
    lib = ctypes.CDLL(rt.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in rest_b200.h but not exported: {missing}"
    # and the ctypes signature table covers the header
    assert sorted(rt.SIGNATURES) == names


def test_reference_ffi_symbols_present(rt):
    # exactly the seven symbols of reference src/external_libs/ffi_restmatr.rs:4-62
    for n in ["ri_ao2mo_f_", "general_dgemm_f_", "special_dgemm_f_01_", "copy_mm_", "copy_mr_", "copy_rm_", "copy_rr_"]:
        assert hasattr(rt.lib, n)


def test_product_does_not_touch_the_oracle():
    pkg = os.path.join(ROOT, "rest_tensors_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".rs")):
                text = open(os.path.join(dirpath, f)).read()
                # functional references only (imports, includes, symbols, the .so); prose in comments is fine
                hits = re.findall(r"import\s+oracle|from\s+oracle|librest_oracle|\borc_[a-z]|#include[^\n]*oracle", text)
                assert not hits, f"{f} uses the oracle: {hits}"


def test_fails_loudly_without_gpu(rt):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(rt.RestB200Error, match="no CPU fallback"):
        rt.MatrixFull.from_vec([2, 2], [1.0, 2.0, 3.0, 4.0]).to_matrixupper()
    from rest_tensors_b200.device import Context
    with pytest.raises(rt.RestB200Error):
        Context(0)


def test_constructors_and_panics(rt):
    r = rt.RIFull.new([3, 2, 2], 1.5)
    assert r.size == [3, 2, 2] and r.indicing == [1, 3, 6] and r.data.size == 12 and np.all(r.data == 1.5)
    with pytest.raises(ValueError):
        rt.RIFull.from_vec([3, 2, 2], np.zeros(11))        # ri.rs:61-63 panic
    assert rt.RIFull.from_vec([3, 2, 2], np.zeros(13)).data.size == 13   # surplus kept (warning in the reference)
    with pytest.raises(ValueError):
        rt.MatrixFull.from_vec([3, 3], np.zeros(8))
    with pytest.raises(ValueError):
        rt.MatrixUpper.from_vec(7, np.zeros(6))
    with pytest.raises(ValueError):
        rt.MatrixFull.new([2, 3], 0.0).to_matrixupper()    # matrixfull.rs:639-641 panic
    assert rt.MatrixUpper.from_vec(4, np.zeros(4)).to_matrixfull() is None   # not triangular -> None
    e = rt.MatrixUpper.empty().to_matrixfull()
    assert e.size == [0, 0] and e.data.size == 0           # matrixupper.rs:336-338
    assert rt.RIFull.empty().size == [0, 0, 0]


def test_index_maps(rt, oracle):
    mu = rt.MatrixUpper.new(10, 0.0)
    assert mu.size2d() == [4, 4]
    for i in range(5):
        for j in range(5):
            assert mu.index2d([i, j]) == oracle.index2d(i, j, 10)
    assert mu.index2d_uncheck([2, 1]) == 3   # no swap: (1+1)*1/2 + 2
    assert mu.index2d([3, 3]) == 9 and mu.index2d([4, 4]) is None


def test_slab_views_are_zero_copy(rt):
    r = rt.RIFull.from_vec([3, 2, 4], np.arange(24.0))
    m = r.get_reducing_matrix(2)
    assert m.size == [3, 2] and m.data.tolist() == list(np.arange(12.0, 18.0))
    m.data[0] = -1.0
    assert r.data[12] == -1.0
    chunks = list(r.iter_auxbas((1, 3)))
    assert len(chunks) == 2 and chunks[0][1] == 7.0
    sub = r.get_reducing_matrix_columns((1, 2), 3)
    assert sub.size == [3, 1] and sub.indicing == [1, 1] and sub.data.tolist() == [21.0, 22.0, 23.0]
    # get_slices: x-runs, z outer / y inner (ri.rs:117-128)
    s = r.get_slices((1, 3), (0, 2), (1, 2))
    assert s.tolist() == [7.0, 8.0, 10.0, 11.0]
    q = rt.MatrixFull.from_vec([2, 6], np.arange(12.0)).to_rifull(2, 3, 2)
    assert q.indicing == [1, 2, 3]   # the reference's quirk (matrixfull.rs:681-685)


def test_views_iterators_and_index_maps(rt, oracle, golden):
    """Host-side index logic of the views (no compute): iter_matrixupper order pinned by the reference's own assert
    (GV4, matrix/mod.rs:437-452), MatrixUpperStepBy, map_upper_to_full / map_full_to_upper, get_slices_mut order."""
    g = golden["GV4"]
    n = g["n"]
    a = rt.MatrixFull.from_vec([n, n], np.array(g["full"], dtype=np.float64))
    assert list(a.iter_matrixupper()) == g["packed"]
    assert rt.MatrixFull.new([2, 3], 0.0).iter_matrixupper() is None and rt.MatrixFull.empty().iter_matrixupper() is None
    # the _mut form walks the same positions; writing through them packs in place
    b = rt.MatrixFull.new([n, n], 0.0)
    for k, pos in enumerate(b.iter_matrixupper_mut()):
        b.data[pos] = float(k)
    bm = b.data.reshape((n, n), order="F")
    for j in range(n):
        for i in range(n):
            assert bm[i, j] == (oracle.index2d(i, j, n * (n + 1) // 2) if i <= j else 0.0)
    # index maps
    for nn in (1, 2, 5, 9):
        ln = nn * (nn + 1) // 2
        up = rt.map_upper_to_full(ln)
        assert up.shape == (ln, 2)
        for t, (i, j) in enumerate(up.tolist()):
            assert i <= j and oracle.index2d(i, j, ln) == t
        full = rt.map_full_to_upper([nn, nn]).reshape((nn, nn), order="F")
        for j in range(nn):
            for i in range(nn):
                assert full[i, j] == (oracle.index2d(i, j, ln) if i <= j else 0)
    assert rt.map_upper_to_full(5) is None and rt.map_full_to_upper([2, 3]) is None
    # MatrixUpperStepBy over an arbitrary iterator, with and without a shift
    assert list(rt.MatrixUpperStepBy(iter("abcdefghi"), [3, 3])) == list("adeghi")
    assert list(rt.MatrixUpperStepBy.new_shift(iter(range(9)), [3, 3], 3)) == [3, 4, 6, 7, 8]
    # get_slices_mut: writable x-runs, z outer / y inner
    r = rt.RIFull.from_vec([3, 2, 4], np.arange(24.0))
    runs = r.get_slices_mut((1, 3), (0, 2), (1, 3))
    assert [x.tolist() for x in runs] == [[7.0, 8.0], [10.0, 11.0], [13.0, 14.0], [16.0, 17.0]]
    runs[2][:] = -1.0
    assert r.data[13] == -1.0 and r.data[14] == -1.0
    assert np.concatenate(r.get_slices_mut_v02((0, 3), (1, 2), (0, 1))).tolist() == r.get_slices((0, 3), (1, 2), (0, 1)).tolist()
    # slices share storage with the parent
    m = rt.MatrixFull.from_vec([3, 4], np.arange(12.0))
    sl = m.to_matrixfullslice_columns((1, 3))
    assert sl.size == [3, 2] and sl.indicing == [0, 3] and sl.data.tolist() == list(np.arange(3.0, 9.0))
    mut = m.to_matrixfullslicemut(); mut.data[0] = 99.0
    assert m.data[0] == 99.0 and m.to_matrixfullslice().get_slice_x(1).tolist() == [3.0, 4.0, 5.0]
    # general_check_shape as written in the reference (equal sizes for NN/TT, reversed for TN/NT)
    p, q = rt.MatrixFull.new([2, 3], 0.0), rt.MatrixFull.new([3, 2], 0.0)
    assert rt.general_check_shape(p, p, 'N', 'N') and not rt.general_check_shape(p, q, 'N', 'N')
    assert rt.general_check_shape(p, q, 'T', 'N') and rt.general_check_shape(p, q, 'N', 'T')
    assert not rt.general_check_shape(p, q, 'X', 'N')
    with pytest.raises(ValueError):
        rt._dgemm_nn(p, p)
    with pytest.raises(ValueError):
        rt._dgemm_tn(p, q)


def test_wrapper_panics_before_ffi(rt):
    a = rt.MatrixFull.new([3, 3], 1.0); b = rt.MatrixFull.new([3, 3], 1.0); c = rt.MatrixFull.new([3, 3], 0.0)
    with pytest.raises(ValueError):     # shape mismatch, matrix_blas_lapack.rs:146-153
        rt._dgemm(a, ((0, 2), (0, 3)), 'N', b, ((0, 2), (0, 2)), 'N', c, ((0, 2), (0, 2)), 1.0, 0.0)
    with pytest.raises(ValueError):     # block outside the matrix, :154-164
        rt._dgemm(a, ((2, 4), (0, 2)), 'N', b, ((0, 2), (0, 2)), 'N', c, ((0, 2), (0, 2)), 1.0, 0.0)
    with pytest.raises(ValueError):     # unknown op -> shape check false -> panic
        rt._dgemm(a, ((0, 2), (0, 2)), 'X', b, ((0, 2), (0, 2)), 'N', c, ((0, 2), (0, 2)), 1.0, 0.0)
    with pytest.raises(ValueError):
        rt._dgemm_full(a, 'N', rt.MatrixFull.new([2, 3], 0.0), 'N', c, 1.0, 0.0)
    with pytest.raises(ValueError):
        rt._dsyrk(a, rt.MatrixFull.new([3, 2], 0.0), 'U', 'N', 1.0, 0.0)
    with pytest.raises(ValueError):
        rt._dgemv(a, np.zeros(2), np.zeros(3), 'N', 1.0, 0.0, 1, 1)
    with pytest.raises(ValueError):     # external_libs/mod.rs:118
        rt.matr_copy(a.data, a.size, (0, 2), (0, 2), c.data, c.size, (0, 3), (0, 2))
    r = rt.RIFull.new([3, 3, 2], 0.0)
    with pytest.raises(ValueError):     # external_libs/mod.rs:188
        r.copy_from_ri((0, 2), (0, 2), (0, 1), rt.RIFull.new([3, 3, 2], 1.0), (0, 2), (0, 2), (0, 2))
    with pytest.raises(ValueError):     # external_libs/mod.rs:163
        r.copy_from_matr((0, 2), (0, 2), 0, 0, a, (0, 3), (0, 2))
    with pytest.raises(ValueError):
        r.self_scaled_add(rt.RIFull.new([3, 3, 1], 0.0), 2.0)
    assert a.add(rt.MatrixFull.new([2, 3], 0.0)) is None   # value-returning forms give None


def test_shard_range_partitions_exactly():
    from rest_tensors_b200.device import shard_range
    for naux in (0, 1, 7, 400, 720, 1700, 4800):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(naux, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == naux
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert shard_range(4800, 3, 8) == (1800, 2400)


def test_split_k_planner(rt):
    """Host-side split-K plan of the GEMM (pure arithmetic, no GPU): the cases DESIGN.md quotes and the invariants."""
    plan = rt.lib.rb_gemm_plan_splits
    sms = 148
    assert plan(600 * 1700, 600, 600, 1, 0, sms) == 1            # ao2mo GEMM 1, config C: ~40 000 tiles, no split
    assert plan(16200, 16200, 1700, 1, 1, sms) == 1              # (ia|jb) diagonal block: 8128 tiles
    assert plan(1700, 1700, 32400, 1, 1, sms) == 7               # RPA-type contraction: 105 tiles x 1013 steps -> 5 rounds
    assert plan(264, 264, 21 * 720, 1, 1, sms) in (48, 49)       # config B SYRK: 3 full + 3 thin tiles -> two rounds
    s_c = plan(600, 600, 60 * 1700, 1, 1, sms)                   # config C SYRK: 15 tiles x 3188 steps
    assert 20 <= s_c <= 40
    assert plan(128, 128, 64, 1, 0, sms) == 1                    # shallow K: never split
    assert plan(128, 128, 32 * 8, 1, 0, sms) == 2                # 8 steps: at most ksteps / 4 splits
    for (m, n, k, batch, tri) in [(100, 100, 10 ** 6, 1, 0), (513, 257, 4099, 3, 0), (2049, 2049, 2049, 1, 2), (64, 8, 10 ** 5, 7, 0)]:
        s = plan(m, n, k, batch, tri, sms)
        ksteps = -(-k // 32)
        assert 1 <= s <= 64 and s <= max(1, ksteps // 4 + 1)
        kper = -(-ksteps // s) * 32
        assert (s - 1) * kper < k                                # every split owns at least one k
        assert s * batch * n * ((m + 1) & ~1) * 8 <= (1 << 30) or s == 1   # partial workspace stays under 1 GB
    assert plan(0, 5, 5, 1, 0, sms) == 1 and plan(5, 5, 5, 0, 0, sms) == 1


def test_p_chunk_planner(rt):
    """P-chunk planner of the RI contractions: whole range when it fits, multiples of 8 otherwise, cheapest GEMM tiling."""
    chunk = rt.lib.rb_ri_plan_chunk
    slab = 1800 * 1800 * 8                                       # config D: W needs nb * nl doubles per slab
    assert chunk(600, slab, 16 << 30, 1) == 600                  # 15.6 GB fits the 16 GB cap: one chunk, flat GEMM 2
    pc = chunk(600, slab, 8 << 30, 1)                            # two chunks: 312 + 288 beats 304 + 296
    assert pc == 312
    for nx, per, budget, tile in [(1700, 600 * 600 * 8, 1 << 30, 1), (4800, slab, 3 << 30, 0), (17, 1 << 20, 1 << 20, 1), (9, 1, 1, 0)]:
        pc = chunk(nx, per, budget, tile)
        assert 1 <= pc <= nx
        assert pc == nx or pc % 8 == 0
        assert pc == nx or pc * per <= max(budget, 8 * per)      # within the budget (at least 8 slabs are always taken)
    assert chunk(0, 8, 8, 0) == 0


def test_argument_errors_surface_before_any_device_call(rt):
    """Shape / range validation of the C ABI and of the host mirror happens before the first CUDA call, so the error
    behaviour (status RB_ERR_INVALID + message; the Rust wrappers panic on the same conditions) is checkable without a GPU."""
    lib, RB_ERR_INVALID = rt.lib, 1
    buf = np.zeros(64)
    p = buf.ctypes.data
    # host-pointer entry points: bad boxes / dimensions
    assert lib.rb_host_ri_iajb(4, p, 2, 2, 0, 3, 0, 2, p, 2, 2, 0, 2, 0, 2, p) == RB_ERR_INVALID       # l range outside
    assert b"box" in lib.rb_last_error()
    assert lib.rb_host_ri_mo_pq(p, 4, 2, 2, 0, 2, 1, 2, None, p) == RB_ERR_INVALID                      # r range outside
    assert lib.rb_host_dspgvx(3, p, p, 4, p, p) == RB_ERR_INVALID                                       # num_orb > n
    assert lib.rb_host_dsyev(b"X", 3, p, p, p) == RB_ERR_INVALID                                        # bad jobz
    assert lib.rb_host_dgemm(b"Q", b"N", 2, 2, 2, 1.0, p, 2, p, 2, 0.0, p, 2) == RB_ERR_INVALID         # bad trans
    assert lib.rb_host_dsyrk(b"U", b"N", 3, 2, 1.0, p, 2, 0.0, p, 3) == RB_ERR_INVALID                  # lda too small
    # device-pointer entry points reject a NULL context first
    assert lib.rb_ri_iajb(None, 4, p, 4, 2, 2, 0, 2, 0, 2, p, 4, 2, 2, 0, 2, 0, 2, 0.0, p, 4) == RB_ERR_INVALID
    assert lib.rb_dsyev(None, b"V", b"L", 3, p, 3, p, p, 3) == RB_ERR_INVALID
    assert lib.rb_special_dgemm_01_peers(None, 0, 1, None, 4, None, None, p, 4, 1.0, 0.0, p) == RB_ERR_INVALID
    # empty problems are not errors and touch nothing
    assert lib.rb_host_ri_iajb(4, p, 2, 2, 0, 0, 0, 2, p, 2, 2, 0, 2, 0, 2, p) == 0
    assert lib.rb_host_dsyev(b"V", 0, p, p, p) == 0 and lib.rb_host_power(0, p, -0.5, 1e-10, p, None) == 0
    # host mirror: the reference's panics
    t = rt.RIFull.new([4, 2, 3], 1.0)
    with pytest.raises(rt.RestB200Error):
        t.ri_iajb((0, 3), (0, 3), (0, 2), (0, 3))
    with pytest.raises(rt.RestB200Error):
        t.ri_iajb((0, 2), (0, 3), (0, 2), (0, 3), other=rt.RIFull.new([5, 2, 3], 1.0))
    with pytest.raises(rt.RestB200Error):
        t.ri_mo_pq((0, 2), (0, 3), np.ones(5))
    with pytest.raises(rt.RestB200Error):
        rt._dsyev(rt.MatrixFull.new([2, 3], 0.0), "V")
    assert rt._power(rt.MatrixFull.new([2, 3], 0.0), -0.5, 1e-10) is None
    assert rt.MatrixFull.new([2, 3], 0.0).lapack_dsyev() is None
    with pytest.raises(rt.RestB200Error):
        rt._dspgvx(rt.MatrixUpper.new(6, 0.0), rt.MatrixUpper.new(10, 0.0), 2)
    with pytest.raises(rt.RestB200Error):
        rt._dspgvx(rt.MatrixUpper.new(6, 0.0), rt.MatrixUpper.new(6, 0.0), 4)


def test_stream_k_partition_covers_every_step_once(rt):
    """Host-side planner of the stream-K GEMM (pure arithmetic, no device call): for a spread of shapes the CTA ranges must tile
    the (tile, k step) space exactly once and in order, no CTA may be empty, and a tile's pieces must be numbered 0 .. pieces-1 by
    consecutive CTAs starting at tile_first (that numbering is what makes the fix-up sum deterministic)."""
    import ctypes as C
    import random
    from rest_tensors_b200._lib import lib
    random.seed(7)
    shapes = [(500, 500, 500, 1, 0), (1000, 1000, 1000, 1, 1), (600, 600, 102000, 1, 1), (264, 264, 15120, 1, 0), (100, 100, 8000, 1, 1),
              (300, 200, 5000, 3, 0), (129, 1030, 777, 1, 0), (1800, 1800, 54720, 1, 1), (2000, 2000, 2000, 1, 0)]
    for _ in range(40):
        tri = random.choice([0, 0, 1, 2])
        m = random.randint(1, 2600)
        n = m if tri else random.randint(1, 2600)
        shapes.append((m, n, random.randint(256, 120000), random.choice([1, 1, 1, 2, 5]), tri))
    checked = 0
    for (m, n, k, batch, tri) in shapes:
        for sms in (148, 132, 7):
            step = (C.c_uint * 161)(); tile = (C.c_ushort * 161)(); first = (C.c_ubyte * 592)(); pieces = (C.c_ubyte * 592)()
            g = lib.rb_gemm_stream_k_tables(m, n, k, batch, tri, sms, C.cast(step, C.c_void_p), C.cast(tile, C.c_void_p),
                                            C.cast(first, C.c_void_p), C.cast(pieces, C.c_void_p))
            if g == 0:
                continue
            checked += 1
            tm, tn = -(-m // 128), -(-n // 128)
            tiles = (tm * (tm + 1) // 2 if tri else tm * tn) * batch
            ksteps = -(-k // 32)
            assert 1 <= g <= sms
            assert (tile[0], step[0]) == (0, 0) and (tile[g], step[g]) == (tiles, 0)
            seen = [0] * tiles            # steps covered so far per tile
            count = [0] * tiles
            for c in range(g):
                t, s = tile[c], step[c]
                te, se = tile[c + 1], step[c + 1]
                assert (t, s) < (te, se), "empty CTA range"
                while (t, s) < (te, se):
                    s1 = se if t == te else ksteps
                    assert seen[t] == s, "steps of a tile must be taken in order, without gaps"
                    assert c - first[t] == count[t], "piece index = CTA - tile_first must count 0, 1, 2, ..."
                    seen[t] = s1; count[t] += 1
                    t, s = t + 1, 0
            assert seen == [ksteps] * tiles and count == list(pieces[:tiles])
    assert checked >= 30
