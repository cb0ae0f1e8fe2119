"""Python (ctypes) handle on the CPU oracle, oracle/librest_oracle.so.

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs -- never from the rest_tensors_b200 package (the product has no CPU path).
See the header of rest_oracle.c for what is restated, from which reference lines, and which parts are
"parity unpinned".
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "librest_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "rest_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


def find_openblas() -> str | None:
    """The LP64 OpenBLAS bundled with scipy (symbols scipy_dgemm_ ...), else a system libopenblas."""
    try:
        import scipy
        hits = glob.glob(os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs", "libscipy_openblas-*.so"))
        if hits:
            return os.path.abspath(hits[0])
    except Exception:
        pass
    for p in ("/usr/lib/x86_64-linux-gnu/libopenblas.so.0", "/usr/lib/x86_64-linux-gnu/libopenblas.so",
              "/usr/lib64/libopenblas.so"):
        if os.path.exists(p):
            return p
    return None


_vp, _i, _i64, _d, _u64 = C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_uint64
_ch = C.c_char


class Oracle:
    def __init__(self):
        self.lib = C.CDLL(build())
        L = self.lib
        L.orc_synth.restype = _d
        L.orc_synth.argtypes = [_u64, _u64, _d]
        L.orc_fill_linear.argtypes = [_vp, _i64, _u64, _u64, _d]
        L.orc_fill_ri3ao_symm.argtypes = [_vp, _i64, _i64, _i64, _u64, _d]
        L.orc_load_blas.restype = _i
        L.orc_load_blas.argtypes = [C.c_char_p]
        L.orc_use_blas.argtypes = [_i]
        L.orc_blas_loaded.restype = _i
        L.orc_blas_set_threads.argtypes = [_i]
        L.orc_blas_get_threads.restype = _i
        L.orc_blas_config.restype = C.c_char_p
        L.orc_dgemm.argtypes = [_ch, _ch, _i, _i, _i, _d, _vp, _i, _vp, _i, _d, _vp, _i]
        L.orc_dsyrk.argtypes = [_ch, _ch, _i, _i, _d, _vp, _i, _d, _vp, _i]
        L.orc_dgemv.argtypes = [_ch, _i, _i, _d, _vp, _i, _vp, _i, _d, _vp, _i]
        L.orc_dsymm.argtypes = [_ch, _ch, _i, _i, _d, _vp, _i, _vp, _i, _d, _vp, _i]
        L.orc_ri_ao2mo_f.argtypes = [_vp, _vp, _vp, _i, _i, _i]
        L.orc_ri_ao2mo_rect.argtypes = [_vp, _i, _vp, _i, _vp, _vp, _i, _i]
        L.orc_ao2mo_v01.argtypes = [_vp, _vp, _vp, _i, _i, _i]
        L.orc_general_dgemm_f.argtypes = [_vp, _i, _i, _i, _i, _i, _i, _ch, _vp, _i, _i, _i, _i, _i, _i, _ch,
                                          _vp, _i, _i, _i, _i, _i, _i, _d, _d]
        L.orc_special_dgemm_f_01.argtypes = [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _i, _i, _i, _i, _i, _d, _d]
        L.orc_copy_mm.argtypes = [_i, _i, _vp, _i, _i, _i, _i, _vp, _i, _i, _i, _i]
        L.orc_copy_mr.argtypes = [_i, _i, _vp, _i, _i, _i, _i, _vp, _i, _i, _i, _i, _i, _i, _i]
        L.orc_copy_rm.argtypes = [_i, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _i, _i, _i]
        L.orc_copy_rr.argtypes = [_i, _i, _i, _vp, _i, _i, _i, _i, _i, _i, _vp, _i, _i, _i, _i, _i, _i]
        L.orc_matrix_transpose.argtypes = [_vp, _i64, _i64, _vp]
        L.orc_ri_transpose.argtypes = [_vp, _i64, _i64, _i64, _i, _vp]
        L.orc_to_matrixupper.argtypes = [_vp, _i64, _vp]
        L.orc_matrixupper_dim.restype = _i64
        L.orc_matrixupper_dim.argtypes = [_i64]
        L.orc_to_matrixfull.restype = _i
        L.orc_to_matrixfull.argtypes = [_vp, _i64, _vp]
        L.orc_index2d.restype = _i64
        L.orc_index2d.argtypes = [_i64, _i64, _i64]
        L.orc_rifull_to_matfull_symm.argtypes = [_vp, _i64, _i64, _vp]
        L.orc_self_scaled_add.argtypes = [_vp, _vp, _d, _i64]
        L.orc_self_general_add.argtypes = [_vp, _vp, _d, _d, _i64]
        L.orc_self_multiple.argtypes = [_vp, _d, _i64]
        L.orc_self_add.argtypes = [_vp, _vp, _i64]
        L.orc_self_sub.argtypes = [_vp, _vp, _i64]
        L.orc_ri_dp.argtypes = [_vp, _vp, _vp, _i, _i]
        L.orc_ri_j.argtypes = [_vp, _vp, _vp, _i, _i]
        L.orc_ri_k.argtypes = [_vp, _vp, _vp, _i, _i, _i]
        L.orc_dsyev.argtypes = [_ch, _i, _vp, _vp]
        L.orc_dspevx.argtypes = [_i, _vp, _vp, _vp, _vp]
        L.orc_dspgvx.argtypes = [_i, _vp, _vp, _i, _vp, _vp, _vp]
        L.orc_power.argtypes = [_vp, _i, _d, _d, _vp, _vp]
        L.orc_ri_mo_pq.argtypes = [_vp, _i, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp]
        L.orc_ri_iajb.argtypes = [_i, _vp, _i, _i, _i, _i, _i, _vp, _i, _i, _i, _i, _i, _vp]
        L.orc_erifold4_chunk_copy_local.restype = _i
        L.orc_erifold4_chunk_copy_local.argtypes = [_vp, _i64, _i64, _i64] + [_i64] * 8 + [_vp]
        L.orc_erifold4_chunk_copy_full.restype = _i
        L.orc_erifold4_chunk_copy_full.argtypes = [_vp, _i64, _i64] + [_i64] * 8 + [_vp]
        for name in ("orc_einsum_01", "orc_einsum_02", "orc_einsum_03"):
            getattr(L, name).argtypes = [_vp, _vp, _vp, _i64, _i64]
            getattr(L, name).restype = None
        self.blas_path = None

    # -- BLAS selection --
    def load_openblas(self, threads: int | None = None) -> bool:
        path = find_openblas()
        if path is None:
            return False
        if self.lib.orc_load_blas(path.encode()) != 0:
            return False
        self.blas_path = path
        if threads:
            self.lib.orc_blas_set_threads(int(threads))
        return True

    def use_blas(self, on: bool) -> None:
        self.lib.orc_use_blas(1 if on else 0)

    def blas_config(self) -> str:
        return self.lib.orc_blas_config().decode()

    def blas_threads(self) -> int:
        return int(self.lib.orc_blas_get_threads())

    def set_threads(self, n: int) -> None:
        self.lib.orc_blas_set_threads(int(n))

    # -- generators --
    def fill_linear(self, n, seed, idx0=0, scale=1.0) -> np.ndarray:
        v = np.empty(int(n), dtype=np.float64)
        self.lib.orc_fill_linear(v.ctypes.data, int(n), seed, idx0, scale)
        return v

    def fill_ri3ao_symm(self, nb, p_lo, p_hi, seed=1, scale=1.0) -> np.ndarray:
        v = np.empty(int(nb) * int(nb) * int(p_hi - p_lo), dtype=np.float64)
        self.lib.orc_fill_ri3ao_symm(v.ctypes.data, nb, p_lo, p_hi, seed, scale)
        return v

    # -- hot path --
    def ri_ao2mo_f(self, c, ri3fn, ns, nb, nx) -> np.ndarray:
        out = np.empty(nx * ns * ns, dtype=np.float64)
        self.lib.orc_ri_ao2mo_f(c.ctypes.data, ri3fn.ctypes.data, out.ctypes.data, ns, nb, nx)
        return out

    def ri_ao2mo_rect(self, cl, nl, cr, nr, ri3fn, nb, nx) -> np.ndarray:
        out = np.empty(nx * nl * nr, dtype=np.float64)
        self.lib.orc_ri_ao2mo_rect(cl.ctypes.data, nl, cr.ctypes.data, nr, ri3fn.ctypes.data, out.ctypes.data, nb, nx)
        return out

    def ao2mo_v01(self, c, ri3fn, ns, nb, nx) -> np.ndarray:
        out = np.empty(nx * nb * ns, dtype=np.float64)
        self.lib.orc_ao2mo_v01(c.ctypes.data, ri3fn.ctypes.data, out.ctypes.data, ns, nb, nx)
        return out

    def ri_dp(self, ri3ao, dm, nb, nx) -> np.ndarray:
        d = np.empty(nx, dtype=np.float64)
        self.lib.orc_ri_dp(ri3ao.ctypes.data, dm.ctypes.data, d.ctypes.data, nb, nx)
        return d

    def ri_j(self, ri3ao, d, nb, nx) -> np.ndarray:
        j = np.empty(nb * nb, dtype=np.float64)
        self.lib.orc_ri_j(ri3ao.ctypes.data, d.ctypes.data, j.ctypes.data, nb, nx)
        return j

    def ri_k(self, ri3ao, ct, nb, no, nx) -> np.ndarray:
        k = np.empty(nb * nb, dtype=np.float64)
        self.lib.orc_ri_k(ri3ao.ctypes.data, ct.ctypes.data, k.ctypes.data, nb, no, nx)
        return k

    def ri_iajb(self, np_, mo_a, nl_a, box_a, mo_b, nl_b, box_b) -> np.ndarray:
        """box = (l0, ll, r0, rl); returns the dense column-major [ll_a*rl_a, ll_b*rl_b] block, flattened"""
        out = np.zeros(box_a[1] * box_a[3] * box_b[1] * box_b[3], dtype=np.float64)
        self.lib.orc_ri_iajb(np_, mo_a.ctypes.data, nl_a, *box_a, mo_b.ctypes.data, nl_b, *box_b, out.ctypes.data)
        return out

    def ri_mo_pq(self, mo_a, npa, mo_b, npb, nl, box, w=None) -> np.ndarray:
        """box = (l0, ll, r0, rl); w = weights over the box's (l, r) pairs or None; returns [npa, npb] flattened"""
        out = np.zeros(npa * npb, dtype=np.float64)
        self.lib.orc_ri_mo_pq(mo_a.ctypes.data, npa, mo_b.ctypes.data, npb, nl, *box,
                              None if w is None else w.ctypes.data, out.ctypes.data)
        return out

    # -- eigen-solvers: the reference's LAPACK calls (needs load_openblas) --
    def dsyev(self, a, n, jobz="V"):
        """returns (eigenvectors [n*n] or None, eigenvalues ascending)"""
        v = np.array(a, dtype=np.float64, copy=True)
        w = np.zeros(n, dtype=np.float64)
        info = self.lib.orc_dsyev(jobz.encode(), n, v.ctypes.data, w.ctypes.data)
        assert info == 0, f"dsyev info {info}"
        return (v if jobz == "V" else None), w

    def dspevx(self, ap, n):
        w = np.zeros(n, dtype=np.float64); z = np.zeros(n * n, dtype=np.float64)
        m = C.c_int(0)
        info = self.lib.orc_dspevx(n, ap.ctypes.data, w.ctypes.data, z.ctypes.data, C.addressof(m))
        assert info == 0, f"dspevx info {info}"
        return z, w, m.value

    def dspgvx(self, ap, bp, n, num_orb):
        w = np.zeros(num_orb, dtype=np.float64); z = np.zeros(n * num_orb, dtype=np.float64)
        m = C.c_int(0)
        info = self.lib.orc_dspgvx(n, ap.ctypes.data, bp.ctypes.data, num_orb, w.ctypes.data, z.ctypes.data, C.addressof(m))
        assert info == 0 and m.value == num_orb, f"dspgvx info {info} m {m.value}"
        return z, w

    def power(self, a, n, p, threshold):
        out = np.zeros(n * n, dtype=np.float64)
        nns = C.c_int(0)
        info = self.lib.orc_power(a.ctypes.data, n, p, threshold, out.ctypes.data, C.addressof(nns))
        assert info == 0, f"power info {info}"
        return out, nns.value

    # -- einsum helpers (matrix_blas_lapack.rs:1273-1387) --
    def einsum_01(self, a, b, ni, nj) -> np.ndarray:
        out = np.empty(ni * nj, dtype=np.float64)
        self.lib.orc_einsum_01(a.ctypes.data, b.ctypes.data, out.ctypes.data, ni, nj)
        return out

    def einsum_02(self, a, b, ni, np_) -> np.ndarray:
        out = np.empty(np_, dtype=np.float64)
        self.lib.orc_einsum_02(a.ctypes.data, b.ctypes.data, out.ctypes.data, ni, np_)
        return out

    def einsum_03(self, a, b, ni, nj) -> np.ndarray:
        out = np.empty(ni * nj, dtype=np.float64)
        self.lib.orc_einsum_03(a.ctypes.data, b.ctypes.data, out.ctypes.data, ni, nj)
        return out

    # -- ERIFold4 chunk copies (src/eri.rs:266-372); return False where the reference would panic --
    def erifold4_chunk_copy_local(self, eri, size, dim, r, buf) -> bool:
        (i0, i1), (j0, j1), (k0, k1), (l0, l1) = r
        return self.lib.orc_erifold4_chunk_copy_local(eri.ctypes.data, size[0], size[1], dim, i0, i1 - i0, j0, j1 - j0, k0, k1 - k0,
                                                      l0, l1 - l0, buf.ctypes.data) == 0

    def erifold4_chunk_copy_full(self, eri, size, r, buf) -> bool:
        (i0, i1), (j0, j1), (k0, k1), (l0, l1) = r
        return self.lib.orc_erifold4_chunk_copy_full(eri.ctypes.data, size[0], size[1], i0, i1 - i0, j0, j1 - j0, k0, k1 - k0, l0,
                                                     l1 - l0, buf.ctypes.data) == 0

    # -- BLAS --
    def dgemm(self, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc) -> None:
        self.lib.orc_dgemm(ta.encode(), tb.encode(), m, n, k, alpha, a.ctypes.data, lda, b.ctypes.data, ldb, beta,
                           c.ctypes.data, ldc)

    def dsyrk(self, uplo, trans, n, k, alpha, a, lda, beta, c, ldc) -> None:
        self.lib.orc_dsyrk(uplo.encode(), trans.encode(), n, k, alpha, a.ctypes.data, lda, beta, c.ctypes.data, ldc)

    def dgemv(self, trans, m, n, alpha, a, lda, x, incx, beta, y, incy) -> None:
        self.lib.orc_dgemv(trans.encode(), m, n, alpha, a.ctypes.data, lda, x.ctypes.data, incx, beta, y.ctypes.data, incy)

    def dsymm(self, side, uplo, m, n, alpha, a, lda, b, ldb, beta, c, ldc) -> None:
        self.lib.orc_dsymm(side.encode(), uplo.encode(), m, n, alpha, a.ctypes.data, lda, b.ctypes.data, ldb, beta,
                           c.ctypes.data, ldc)

    def general_dgemm_f(self, a, size_a, ra, ca, opa, b, size_b, rb, cb, opb, c, size_c, rc, cc, alpha, beta) -> None:
        self.lib.orc_general_dgemm_f(a.ctypes.data, size_a[0], size_a[1], ra[0], ra[1] - ra[0], ca[0], ca[1] - ca[0],
                                     opa.encode(), b.ctypes.data, size_b[0], size_b[1], rb[0], rb[1] - rb[0], cb[0],
                                     cb[1] - cb[0], opb.encode(), c.ctypes.data, size_c[0], size_c[1], rc[0],
                                     rc[1] - rc[0], cc[0], cc[1] - cc[0], alpha, beta)

    def special_dgemm_f_01(self, t, size_a, rx, i_y, rz, b, size_b, rrb, rcb, alpha, beta) -> None:
        self.lib.orc_special_dgemm_f_01(t.ctypes.data, size_a[0], size_a[1], size_a[2], rx[0], rx[1] - rx[0], i_y, rz[0],
                                        rz[1] - rz[0], b.ctypes.data, size_b[0], size_b[1], rrb[0], rrb[1] - rrb[0],
                                        rcb[0], rcb[1] - rcb[0], alpha, beta)

    # -- copies --
    def copy_mm(self, xl, yl, f, fx, fy, fxs, fys, t, tx, ty, txs, tys) -> None:
        self.lib.orc_copy_mm(xl, yl, f.ctypes.data, fx, fy, fxs, fys, t.ctypes.data, tx, ty, txs, tys)

    def copy_mr(self, xl, yl, f, fx, fy, fxs, fys, t, tx, ty, tz, txs, tys, t3, mod) -> None:
        self.lib.orc_copy_mr(xl, yl, f.ctypes.data, fx, fy, fxs, fys, t.ctypes.data, tx, ty, tz, txs, tys, t3, mod)

    def copy_rm(self, xl, yl, f, fx, fy, fz, fxs, fys, f3, mod, t, tx, ty, txs, tys) -> None:
        self.lib.orc_copy_rm(xl, yl, f.ctypes.data, fx, fy, fz, fxs, fys, f3, mod, t.ctypes.data, tx, ty, txs, tys)

    def copy_rr(self, xl, yl, zl, f, fx, fy, fz, fxs, fys, fzs, t, tx, ty, tz, txs, tys, tzs) -> None:
        self.lib.orc_copy_rr(xl, yl, zl, f.ctypes.data, fx, fy, fz, fxs, fys, fzs, t.ctypes.data, tx, ty, tz, txs, tys, tzs)

    # -- layout --
    def matrix_transpose(self, a, rows, cols) -> np.ndarray:
        out = np.empty(rows * cols, dtype=np.float64)
        self.lib.orc_matrix_transpose(a.ctypes.data, rows, cols, out.ctypes.data)
        return out

    def ri_transpose(self, a, i, j, k, which) -> np.ndarray:
        out = np.empty(i * j * k, dtype=np.float64)
        self.lib.orc_ri_transpose(a.ctypes.data, i, j, k, which, out.ctypes.data)
        return out

    def to_matrixupper(self, full, n) -> np.ndarray:
        out = np.empty(n * (n + 1) // 2, dtype=np.float64)
        self.lib.orc_to_matrixupper(full.ctypes.data, n, out.ctypes.data)
        return out

    def to_matrixfull(self, packed):
        n = self.lib.orc_matrixupper_dim(packed.size)
        if n < 0:
            return None
        out = np.empty(n * n, dtype=np.float64)
        self.lib.orc_to_matrixfull(packed.ctypes.data, packed.size, out.ctypes.data)
        return out

    def index2d(self, i, j, ln):
        r = self.lib.orc_index2d(i, j, ln)
        return None if r < 0 else int(r)

    def rifull_to_matfull_symm(self, ri, nao, naux) -> np.ndarray:
        out = np.empty(nao * (nao + 1) // 2 * naux, dtype=np.float64)
        self.lib.orc_rifull_to_matfull_symm(ri.ctypes.data, nao, naux, out.ctypes.data)
        return out

    def axpy(self, op, c, p, a, b) -> None:
        n = c.size
        if op == 0:
            self.lib.orc_self_scaled_add(c.ctypes.data, p.ctypes.data, b, n)
        elif op == 1:
            self.lib.orc_self_general_add(c.ctypes.data, p.ctypes.data, a, b, n)
        elif op == 2:
            self.lib.orc_self_multiple(c.ctypes.data, a, n)
        elif op == 3:
            self.lib.orc_self_add(c.ctypes.data, p.ctypes.data, n)
        else:
            self.lib.orc_self_sub(c.ctypes.data, p.ctypes.data, n)
