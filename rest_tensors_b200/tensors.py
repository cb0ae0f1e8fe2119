"""Host-side mirror of the reference's tensor API for the RI hot path.

Same type names, field names (``size`` / ``indicing`` / ``data``), method names, argument meaning and error
behaviour as the reference's Rust structs, so that the parity tests read like the reference's own tests:

    RIFull        src/ri.rs:18-433
    MatrixFull    src/matrix/mod.rs:472-480, src/matrix/matrixfull.rs
    MatrixUpper   src/matrix/matrixupper.rs:231-420, src/index.rs:209-233
    MatrixFullSlice(Mut), MatrixUpperSlice, ERIFold4, MatrixUpperStepBy, map_upper_to_full, map_full_to_upper
    _dgemm, _dgemm_full, _dgemm_full_new, _dsyrk, _dsymm, _dgemv, general_check_shape,
    _dgemm_nn(_serial), _dgemm_tn(_serial), _dgemm_tn_v02            src/matrix/matrix_blas_lapack.rs
    ri_ao2mo_f, general_dgemm_f, special_dgemm_f_01, matr_copy, ...   src/external_libs/mod.rs

``data`` is a flat column-major ``numpy.float64`` array (the Rust ``Vec<f64>``).  A Rust ``panic!`` is a Python
exception here (``ValueError`` for the shape panics, ``RestB200Error`` for library failures); ``Option::None``
is ``None``.  Every numerical operation and every bulk data movement is a call into librest_b200.so; NumPy only
owns the buffers (views, slicing, allocation) exactly as ``Vec``/slices do on the Rust side.  There is no
NumPy/CPU implementation of any operation behind these methods.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterator, Optional, Sequence, Tuple

import numpy as np

from ._lib import lib, check, ch, RestB200Error  # noqa: F401

Range = Tuple[int, int]  # half-open (start, end), the Rust `start..end`


def _ptr(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)


def _f64(v, n: Optional[int] = None) -> np.ndarray:
    a = np.ascontiguousarray(np.asarray(v, dtype=np.float64).reshape(-1))
    return a


def _ci(v: int) -> C.c_int:
    return C.c_int(int(v))


def _rlen(r: Range) -> int:
    return max(0, int(r[1]) - int(r[0]))


def _indicing(size: Sequence[int]):
    ind, ln = [], 1
    for s in size:
        ind.append(ln)
        ln *= int(s)
    return ind, ln


# ======================================================================================================
# external_libs (src/external_libs/mod.rs) -- safe wrappers over the Fortran-ABI symbols
# ======================================================================================================
def ri_ao2mo_f(eigenvector: np.ndarray, ri3fn: np.ndarray, ri3mo: np.ndarray, num_states: int, num_basis: int,
               num_auxbas: int) -> None:
    """src/external_libs/mod.rs:6-22 -> ffi ri_ao2mo_f_"""
    lib.ri_ao2mo_f_(_ptr(eigenvector), _ptr(ri3fn), _ptr(ri3mo), C.byref(_ci(num_states)), C.byref(_ci(num_basis)),
                    C.byref(_ci(num_auxbas)))


def general_dgemm_f(matr_a, size_a, range_row_a: Range, range_column_a: Range, opa: str,
                    matr_b, size_b, range_row_b: Range, range_column_b: Range, opb: str,
                    matr_c, size_c, range_row_c: Range, range_column_c: Range, alpha: float, beta: float) -> None:
    """src/external_libs/mod.rs:60-83 -> ffi general_dgemm_f_"""
    a = [_ci(size_a[0]), _ci(size_a[1]), _ci(range_row_a[0]), _ci(_rlen(range_row_a)), _ci(range_column_a[0]),
         _ci(_rlen(range_column_a))]
    b = [_ci(size_b[0]), _ci(size_b[1]), _ci(range_row_b[0]), _ci(_rlen(range_row_b)), _ci(range_column_b[0]),
         _ci(_rlen(range_column_b))]
    c = [_ci(size_c[0]), _ci(size_c[1]), _ci(range_row_c[0]), _ci(_rlen(range_row_c)), _ci(range_column_c[0]),
         _ci(_rlen(range_column_c))]
    al, be = C.c_double(alpha), C.c_double(beta)
    lib.general_dgemm_f_(_ptr(matr_a), *[C.byref(v) for v in a], ch(opa), _ptr(matr_b), *[C.byref(v) for v in b], ch(opb),
                         _ptr(matr_c), *[C.byref(v) for v in c], C.byref(al), C.byref(be))


def special_dgemm_f_01(ten3_a, size_a, range_x_a: Range, i_y: int, range_z_a: Range, matr_b, size_b,
                       range_row_b: Range, range_column_b: Range, alpha: float, beta: float) -> None:
    """src/external_libs/mod.rs:85-99 -> ffi special_dgemm_f_01_"""
    v = [_ci(size_a[0]), _ci(size_a[1]), _ci(size_a[2]), _ci(range_x_a[0]), _ci(_rlen(range_x_a)), _ci(i_y),
         _ci(range_z_a[0]), _ci(_rlen(range_z_a))]
    w = [_ci(size_b[0]), _ci(size_b[1]), _ci(range_row_b[0]), _ci(_rlen(range_row_b)), _ci(range_column_b[0]),
         _ci(_rlen(range_column_b))]
    al, be = C.c_double(alpha), C.c_double(beta)
    lib.special_dgemm_f_01_(_ptr(ten3_a), *[C.byref(x) for x in v], _ptr(matr_b), *[C.byref(x) for x in w],
                            C.byref(al), C.byref(be))


def matr_copy(matr_a, size_a, range_row_a: Range, range_column_a: Range, matr_b, size_b, range_row_b: Range,
              range_column_b: Range) -> None:
    """src/external_libs/mod.rs:102-120: from matr_a[(ra, ca)] to matr_b[(rb, cb)]"""
    x_len, y_len = _rlen(range_row_a), _rlen(range_column_a)
    if not (x_len == _rlen(range_row_b) and y_len == _rlen(range_column_b)):
        raise ValueError("Error: the data block for copy has different size between two matrices")
    v = [_ci(x_len), _ci(y_len)]
    f = [_ci(size_a[0]), _ci(size_a[1]), _ci(range_row_a[0]), _ci(range_column_a[0])]
    t = [_ci(size_b[0]), _ci(size_b[1]), _ci(range_row_b[0]), _ci(range_column_b[0])]
    lib.copy_mm_(*[C.byref(x) for x in v], _ptr(matr_a), *[C.byref(x) for x in f], _ptr(matr_b),
                 *[C.byref(x) for x in t])


def matr_copy_from_ri(ri_a, size_a, range_x_a: Range, range_y_a: Range, i_z_a: int, copy_mod: int, matr_b, size_b,
                      range_row_b: Range, range_column_b: Range) -> None:
    """src/external_libs/mod.rs:123-140 -> copy_rm_"""
    x_len, y_len = _rlen(range_x_a), _rlen(range_y_a)
    if not (x_len == _rlen(range_row_b) and y_len == _rlen(range_column_b)):
        raise ValueError("Error: the data block for copy has different size between matrix and ri-tensor")
    v = [_ci(x_len), _ci(y_len)]
    f = [_ci(size_a[0]), _ci(size_a[1]), _ci(size_a[2]), _ci(range_x_a[0]), _ci(range_y_a[0]), _ci(i_z_a), _ci(copy_mod)]
    t = [_ci(size_b[0]), _ci(size_b[1]), _ci(range_row_b[0]), _ci(range_column_b[0])]
    lib.copy_rm_(*[C.byref(x) for x in v], _ptr(ri_a), *[C.byref(x) for x in f], _ptr(matr_b),
                 *[C.byref(x) for x in t])


def ri_copy_from_matr(matr_a, size_a, range_row_a: Range, range_column_a: Range, ri_b, size_b, range_row_b: Range,
                      range_column_b: Range, i_high_b: int, copy_mod: int) -> None:
    """src/external_libs/mod.rs:145-165 -> copy_mr_"""
    x_len, y_len = _rlen(range_row_a), _rlen(range_column_a)
    if not (x_len == _rlen(range_row_b) and y_len == _rlen(range_column_b)):
        raise ValueError("Error: the data block for copy has different size between the matrix and ri 3D-tensor")
    v = [_ci(x_len), _ci(y_len)]
    f = [_ci(size_a[0]), _ci(size_a[1]), _ci(range_row_a[0]), _ci(range_column_a[0])]
    t = [_ci(size_b[0]), _ci(size_b[1]), _ci(size_b[2]), _ci(range_row_b[0]), _ci(range_column_b[0]), _ci(i_high_b),
         _ci(copy_mod)]
    lib.copy_mr_(*[C.byref(x) for x in v], _ptr(matr_a), *[C.byref(x) for x in f], _ptr(ri_b),
                 *[C.byref(x) for x in t])


def ri_copy_from_ri(ri_a, size_a, range_x_a: Range, range_y_a: Range, range_z_a: Range, ri_b, size_b,
                    range_x_b: Range, range_y_b: Range, range_z_b: Range) -> None:
    """src/external_libs/mod.rs:168-190 -> copy_rr_"""
    x_len, y_len, z_len = _rlen(range_x_a), _rlen(range_y_a), _rlen(range_z_a)
    if not (x_len == _rlen(range_x_b) and y_len == _rlen(range_y_b) and z_len == _rlen(range_z_b)):
        raise ValueError("Error: the data block for copy has different size between ri 3D-tensors")
    v = [_ci(x_len), _ci(y_len), _ci(z_len)]
    f = [_ci(size_a[0]), _ci(size_a[1]), _ci(size_a[2]), _ci(range_x_a[0]), _ci(range_y_a[0]), _ci(range_z_a[0])]
    t = [_ci(size_b[0]), _ci(size_b[1]), _ci(size_b[2]), _ci(range_x_b[0]), _ci(range_y_b[0]), _ci(range_z_b[0])]
    lib.copy_rr_(*[C.byref(x) for x in v], _ptr(ri_a), *[C.byref(x) for x in f], _ptr(ri_b),
                 *[C.byref(x) for x in t])


# ======================================================================================================
# MatrixFull
# ======================================================================================================
class MatrixFull:
    """src/matrix/mod.rs:472-480: ``size: [usize;2]``, ``indicing: [usize;2]``, ``data: Vec<T>`` (column-major)."""

    def __init__(self, size, indicing, data: np.ndarray):
        self.size = [int(size[0]), int(size[1])]
        self.indicing = list(indicing)
        self.data = data

    # -- constructors (matrixfull.rs:190-240) --
    @staticmethod
    def new(size, new_default: float) -> "MatrixFull":
        ind, ln = _indicing(size)
        return MatrixFull(size, ind, np.full(ln, float(new_default), dtype=np.float64))

    @staticmethod
    def empty() -> "MatrixFull":
        return MatrixFull([0, 0], [0, 0], np.zeros(0, dtype=np.float64))

    @staticmethod
    def from_vec(size, new_vec) -> "MatrixFull":
        ind, ln = _indicing(size)
        data = _f64(new_vec)
        if ln > data.size:
            raise ValueError("Error: inconsistency happens when formating a matrix from a given vector, "
                             f"(length from size, length of new vector) = ({ln},{data.size})")
        return MatrixFull(size, ind, data)

    # -- BasicMatrix (matrix/mod.rs:490-505) --
    def data_ref(self) -> np.ndarray:
        return self.data

    def data_ref_mut(self) -> np.ndarray:
        return self.data

    def is_contiguous(self) -> bool:
        return True

    def check_shape(self, other: "MatrixFull") -> bool:
        return self.size[0] == other.size[0] and self.size[1] == other.size[1]

    def get2d(self, pos) -> float:
        return float(self.data[pos[0] * self.indicing[0] + pos[1] * self.indicing[1]])

    def reshape(self, size) -> None:
        """matrixfull.rs:242-253 (metadata only)"""
        ind, ln = _indicing(size)
        if ln != self.size[0] * self.size[1]:
            raise ValueError("Cannot reshape: element counts differ")
        self.size, self.indicing = [int(size[0]), int(size[1])], ind

    # -- layout ops --
    def transpose(self) -> "MatrixFull":
        """matrixfull.rs:579-596"""
        r, c = self.size
        out = MatrixFull.new([c, r], 0.0)
        if r * c:
            check(lib.rb_host_matrix_transpose(_ptr(self.data), r, c, _ptr(out.data)), "MatrixFull::transpose")
        return out

    transpose_and_drop = transpose

    def to_matrixupper(self) -> "MatrixUpper":
        """matrixfull.rs:638-646: panics unless square; packed[j(j+1)/2+i] = a[i+j*n], i<=j"""
        if self.size[0] != self.size[1]:
            raise ValueError("Error: Nonsymmetric matrix cannot be converted to the upper format")
        n = self.size[0]
        out = np.zeros(n * (n + 1) // 2, dtype=np.float64)
        check(lib.rb_host_to_matrixupper(_ptr(self.data), n, _ptr(out)), "MatrixFull::to_matrixupper")
        return MatrixUpper(out.size, out)

    def to_rifull(self, i: int, j: int, k: int) -> "RIFull":
        """matrixfull.rs:674-686 (note the reference's indicing quirk [1, i, j])"""
        if i * j * k != self.size[0] * self.size[1]:
            raise ValueError("Error in tranforming MatrixFull to RIFull: incompitable size")
        return RIFull([i, j, k], [1, i, j], self.data.copy())

    # -- views (matrixfull.rs:648-672): zero-copy, share ``data`` with the parent like the Rust borrows --
    def to_matrixfullslice(self) -> "MatrixFullSlice":
        return MatrixFullSlice(self.size[0:2], self.indicing[0:2], self.data)

    def to_matrixfullslicemut(self) -> "MatrixFullSliceMut":
        return MatrixFullSliceMut(self.size[0:2], self.indicing[0:2], self.data)

    def to_matrixfullslice_columns(self, range_columns: Range) -> "MatrixFullSlice":
        """matrixfull.rs:664-672 (keeps the reference's indicing quirk [0, rows])"""
        start = range_columns[0] * self.indicing[1]
        end = start + _rlen(range_columns) * self.indicing[1]
        return MatrixFullSlice([self.size[0], _rlen(range_columns)], [0, self.size[0]], self.data[start:end])

    def iter_matrixupper(self) -> Optional["MatrixUpperStepBy"]:
        """matrixfull.rs:407-413: None unless square and non-empty; yields a[i + j*n], i <= j, in packed order"""
        x, y = self.size
        if x == 0 or y == 0 or x != y:
            return None
        return MatrixUpperStepBy(iter(range(x * y)), [x, y], source=self.data)

    def iter_matrixupper_mut(self) -> Optional["MatrixUpperStepBy"]:
        """matrixfull.rs:415-423: same walk; the items are linear positions to assign through (`data[pos] = v`)"""
        x, y = self.size
        if x == 0 or y == 0 or x != y:
            return None
        return MatrixUpperStepBy(iter(range(x * y)), [x, y])

    def copy_from_matr(self, range_x: Range, range_y: Range, from_matr: "MatrixFull", f_range_x: Range,
                       f_range_y: Range) -> None:
        """matrixfull.rs:1388-1396 -> matr_copy -> copy_mm_"""
        matr_copy(from_matr.data, from_matr.size, f_range_x, f_range_y, self.data, self.size, range_x, range_y)

    def lapack_dgemm(self, a: "MatrixFull", b: "MatrixFull", opa: str, opb: str, alpha: float, beta: float) -> None:
        """matrix_blas_lapack.rs:739-774 (no shape check, like the reference)"""
        m = a.size[0] if opa == 'N' else a.size[1]
        k = a.size[1] if opa == 'N' else a.size[0]
        n = b.size[1] if opb == 'N' else b.size[0]
        lda = max(m, 1) if opa == 'N' else max(k, 1)
        ldb = max(k, 1) if opb == 'N' else max(n, 1)
        check(lib.rb_host_dgemm(ch(opa), ch(opb), m, n, k, alpha, _ptr(a.data), lda, _ptr(b.data), ldb, beta,
                                _ptr(self.data), max(m, 1)), "lapack_dgemm")

    def ddot(self, b: "MatrixFull") -> Optional["MatrixFull"]:
        """matrix_blas_lapack.rs:714-730: a*b with beta=1 on a zeroed C"""
        if self.size[1] != b.size[0]:
            return None
        m, n, k = self.size[0], b.size[1], self.size[1]
        c = MatrixFull.new([m, n], 0.0)
        check(lib.rb_host_dgemm(b'N', b'N', m, n, k, 1.0, _ptr(self.data), max(m, 1), _ptr(b.data), max(k, 1), 1.0,
                                _ptr(c.data), max(m, 1)), "ddot")
        return c

    # -- MathMatrix (matrix/mod.rs:545-648) --
    def _axpy(self, op: int, other: Optional["MatrixFull"], a: float, b: float) -> None:
        p = _ptr(other.data) if other is not None else None
        check(lib.rb_host_axpy(op, _ptr(self.data), p, a, b, self.size[0] * self.size[1]), "axpy")

    def add(self, other: "MatrixFull") -> Optional["MatrixFull"]:
        if not self.check_shape(other):
            return None
        out = MatrixFull(self.size, self.indicing, self.data.copy())
        out._axpy(3, other, 0.0, 0.0)
        return out

    def scaled_add(self, other: "MatrixFull", scale_factor: float) -> Optional["MatrixFull"]:
        if not self.check_shape(other):
            return None
        out = MatrixFull(self.size, self.indicing, self.data.copy())
        out._axpy(0, other, 0.0, scale_factor)
        return out

    def sub(self, other: "MatrixFull") -> Optional["MatrixFull"]:
        if not self.check_shape(other):
            return None
        out = MatrixFull(self.size, self.indicing, self.data.copy())
        out._axpy(4, other, 0.0, 0.0)
        return out

    def _need_shape(self, bm: "MatrixFull", what: str) -> None:
        if not self.check_shape(bm):
            raise ValueError(f"Error: Shape inconsistency happens when {what} two matrices")

    def self_add(self, bm: "MatrixFull") -> None:
        self._need_shape(bm, "plus")
        self._axpy(3, bm, 0.0, 0.0)

    def self_sub(self, bm: "MatrixFull") -> None:
        self._need_shape(bm, "subtract")
        self._axpy(4, bm, 0.0, 0.0)

    def self_scaled_add(self, bm: "MatrixFull", b: float) -> None:
        self._need_shape(bm, "plus")
        self._axpy(0, bm, 0.0, b)

    def self_general_add(self, bm: "MatrixFull", a: float, b: float) -> None:
        self._need_shape(bm, "plus")
        self._axpy(1, bm, a, b)

    def self_multiple(self, a: float) -> None:
        self._axpy(2, None, a, 0.0)

    # -- eigen-solvers (matrix_blas_lapack.rs:775-797, 1004-1062): one-sided Jacobi on the GPU behind the LAPACK names --
    def lapack_dsyev(self):
        """(eigenvectors [n, n], eigenvalues ascending, n) or None for a non-square matrix"""
        if self.size[0] != self.size[1]:
            return None
        vec, w, n = _dsyev(self, "V")
        return vec, w, n

    def lapack_power(self, p: float, threshold: float) -> Optional["MatrixFull"]:
        return _power(self, p, threshold)


# ======================================================================================================
# MatrixUpper
# ======================================================================================================
class MatrixUpper:
    """src/matrix/matrixupper.rs:231-234: ``size`` = n(n+1)/2 (the packed length), ``data``."""

    def __init__(self, size: int, data: np.ndarray):
        self.size = int(size)
        self.data = data

    @staticmethod
    def new(size: int, new_default: float) -> "MatrixUpper":
        return MatrixUpper(size, np.full(int(size), float(new_default), dtype=np.float64))

    @staticmethod
    def empty() -> "MatrixUpper":
        return MatrixUpper(0, np.zeros(0, dtype=np.float64))

    @staticmethod
    def from_vec(size: int, new_vec) -> "MatrixUpper":
        data = _f64(new_vec)
        if size > data.size:
            raise ValueError("Error: inconsistency happens when formating a matrix from a given vector, "
                             f"(length from size, length of new vector) = ({size},{data.size})")
        return MatrixUpper(size, data)

    def len(self) -> int:
        return int(self.data.size)

    def size2d(self):
        """matrixupper.rs:289-292 (`size()` in the reference): [n, n] from the packed length"""
        n = int((1.0 + 8.0 * float(self.size)) ** 0.5 * 0.5 - 0.5)
        return [n, n]

    def index2d(self, position) -> Optional[int]:
        """index.rs:209-226: swaps so that i<=j; None when out of range"""
        i, j = (position[0], position[1]) if position[0] <= position[1] else (position[1], position[0])
        tp = (j + 1) * j // 2 + i
        return tp if tp < self.data.size else None

    def index2d_uncheck(self, position) -> Optional[int]:
        """index.rs:227-233: no swap"""
        tp = (position[1] + 1) * position[1] // 2 + position[0]
        return tp if tp < self.data.size else None

    def get2d(self, position) -> Optional[float]:
        idx = self.index2d(position)
        return None if idx is None else float(self.data[idx])

    def to_matrixfull(self) -> Optional[MatrixFull]:
        """matrixupper.rs:330-373: None unless size is triangular; empty() for length 0; unpack + mirror"""
        if self.len() == 0:
            return MatrixFull.empty()
        n = int((1.0 + 8.0 * float(self.size)) ** 0.5 * 0.5 - 0.5)
        if n * (n + 1) // 2 != self.size:
            return None
        out = MatrixFull.new([n, n], 0.0)
        check(lib.rb_host_to_matrixfull(_ptr(self.data), self.size, _ptr(out.data)), "MatrixUpper::to_matrixfull")
        return out

    def to_matrixupperslice(self) -> "MatrixUpperSlice":
        return MatrixUpperSlice(self.data)

    def _zip(self, other: "MatrixUpper", op: int) -> "MatrixUpper":
        """matrixupper.rs:395-420: zip silently truncates to the shorter operand"""
        out = MatrixUpper(self.size, self.data.copy())
        n = min(out.data.size, other.data.size)
        if n:
            check(lib.rb_host_axpy(op, _ptr(out.data), _ptr(other.data), 0.0, 0.0, n), "MatrixUpper add/sub")
        return out

    def __add__(self, other: "MatrixUpper") -> "MatrixUpper":
        return self._zip(other, 3)

    def __sub__(self, other: "MatrixUpper") -> "MatrixUpper":
        return self._zip(other, 4)

    # -- packed eigen-solvers (matrix_blas_lapack.rs:1075-1147) --
    def lapack_dspevx(self):
        """(eigenvectors [n, n], eigenvalues ascending, n_found) of the packed-upper symmetric matrix"""
        n = self.size2d()[0]
        if n * (n + 1) // 2 != self.data.size:
            raise RestB200Error("lapack_dspevx: the packed length is not triangular")
        w = np.zeros(n, dtype=np.float64)
        z = MatrixFull.new([n, n], 0.0)
        found = C.c_int(0)
        check(lib.rb_host_dspevx(n, _ptr(self.data), _ptr(w), _ptr(z.data), C.byref(found)), "lapack_dspevx")
        return z, w, found.value

    def lapack_dspgvx(self, ovlp: "MatrixUpper", num_orb: int):
        """solve A x = lambda B x (A = self, B = ovlp, both packed upper): (eigenvectors [n, num_orb], eigenvalues)"""
        return _dspgvx(self, ovlp, num_orb)


# ======================================================================================================
# RIFull
# ======================================================================================================
class RIFull:
    """src/ri.rs:18-24: rank-3 column-major tensor, linear index x + y*s0 + z*s0*s1."""

    def __init__(self, size, indicing, data: np.ndarray):
        self.size = [int(s) for s in size]
        self.indicing = list(indicing)
        self.data = data

    @staticmethod
    def new(size, new_default: float) -> "RIFull":
        ind, ln = _indicing(size)
        return RIFull(size, ind, np.full(ln, float(new_default), dtype=np.float64))

    @staticmethod
    def empty() -> "RIFull":
        return RIFull([0, 0, 0], [0, 0, 0], np.zeros(0, dtype=np.float64))

    @staticmethod
    def from_vec(size, new_vec) -> "RIFull":
        """ri.rs:57-70: panics when the vector is too short, keeps (and warns about) a surplus"""
        ind, ln = _indicing(size)
        data = _f64(new_vec)
        if ln > data.size:
            raise ValueError("Error: inconsistency happens when formating a tensor from a given vector, "
                             f"(length from size, length of new vector) = ({ln},{data.size})")
        return RIFull(size, ind, data)

    def check_shape(self, other: "RIFull") -> bool:
        return all(a == b for a, b in zip(self.size, other.size))

    # -- zero-copy slab access (ri.rs:71-218): pointer arithmetic only, as in the reference --
    def get_reducing_matrix(self, i_reduced: int) -> MatrixFull:
        p_length = self.indicing[2]
        p_start = p_length * i_reduced
        return MatrixFull(self.size[0:2], self.indicing[0:2], self.data[p_start:p_start + p_length])

    get_reducing_matrix_mut = get_reducing_matrix

    def get_reducing_matrix_columns(self, range_columns: Range, i_reduced: int) -> MatrixFull:
        """ri.rs:102-115 (keeps the reference's indicing quirk [1, |cols|])"""
        z_length, y_length = self.indicing[2], self.indicing[1]
        start = z_length * i_reduced + y_length * range_columns[0]
        end = start + y_length * _rlen(range_columns)
        size = [self.size[0], _rlen(range_columns)]
        return MatrixFull(size, [1, size[1]], self.data[start:end])

    def iter_auxbas(self, auxbas_range: Range) -> Iterator[np.ndarray]:
        """ri.rs:190-198: the P-shard primitive -- contiguous chunks of s0*s1"""
        chunk = self.size[0] * self.size[1]
        for p in range(auxbas_range[0], auxbas_range[1]):
            yield self.data[chunk * p: chunk * (p + 1)]

    iter_mut_auxbas = iter_auxbas
    par_iter_auxbas = iter_auxbas
    par_iter_mut_auxbas = iter_auxbas

    def iter_slices_x(self, y: int, z: int) -> np.ndarray:
        start = z * self.indicing[2] + y * self.indicing[1]
        return self.data[start:start + self.indicing[1]]

    def get_slices(self, x: Range, y: Range, z: Range) -> np.ndarray:
        """ri.rs:117-128: x-runs flattened in (z outer, y inner) order (views into ``data`` concatenated)"""
        s0, s1, s2 = self.size
        cube = self.data[: s0 * s1 * s2].reshape((s0, s1, s2), order="F")
        return cube[x[0]:x[1], y[0]:y[1], z[0]:z[1]].reshape(-1, order="F")

    def get_slices_mut(self, x: Range, y: Range, z: Range):
        """ri.rs:130-166 (`get_slices_mut`, `_v01`, `_v02` differ only in how the Vec is built): the x-runs as a list
        of writable views into ``data`` in (z outer, y inner) order -- the Rust `Vec<&mut [T]>` before flattening"""
        len_y, len_z = self.indicing[1], self.indicing[2]
        out = []
        for zz in range(z[0], z[1]):
            for yy in range(y[0], y[1]):
                start = x[0] + yy * len_y + zz * len_z
                out.append(self.data[start:start + _rlen(x)])
        return out

    get_slices_mut_v01 = get_slices_mut
    get_slices_mut_v02 = get_slices_mut

    # -- transposes (ri.rs:227-294) --
    def _transpose(self, which: int, new_size) -> "RIFull":
        i, j, k = self.size
        out = RIFull.new(new_size, 0.0)
        if i * j * k:
            check(lib.rb_host_ri_transpose(_ptr(self.data), i, j, k, which, _ptr(out.data)), "RIFull::transpose")
        return out

    def transpose_jik(self) -> "RIFull":
        return self._transpose(0, [self.size[1], self.size[0], self.size[2]])

    def transpose_jki(self) -> "RIFull":
        return self._transpose(1, [self.size[1], self.size[2], self.size[0]])

    def transpose_kji(self) -> "RIFull":
        return self._transpose(2, [self.size[2], self.size[1], self.size[0]])

    def transpose_ikj(self) -> "RIFull":
        return self._transpose(3, [self.size[0], self.size[2], self.size[1]])

    # -- reshapes (ri.rs:297-326) --
    def rifull_to_matfull_symm(self) -> MatrixFull:
        nao, naux = self.size[0], self.size[2]
        out = MatrixFull.new([nao * (nao + 1) // 2, naux], 0.0)
        if nao * naux:
            check(lib.rb_host_ri_pack_symm(_ptr(self.data), nao, naux, _ptr(out.data)), "rifull_to_matfull_symm")
        return out

    def rifull_to_matfull_ij_k(self) -> MatrixFull:
        i, j, k = self.size
        return MatrixFull.from_vec([i * j, k], self.data.copy())

    def rifull_to_matfull_i_jk(self) -> MatrixFull:
        i, j, k = self.size
        return MatrixFull.from_vec([i, j * k], self.data.copy())

    # -- arithmetic --
    def self_scaled_add(self, bm: "RIFull", b: float) -> None:
        """ri.rs:345-354: A += b*B"""
        if not self.check_shape(bm):
            raise ValueError("Error: Shape inconsistency happens when plus two matrices")
        n = self.size[0] * self.size[1] * self.size[2]
        check(lib.rb_host_axpy(0, _ptr(self.data), _ptr(bm.data), 0.0, b, n), "RIFull::self_scaled_add")

    # -- the hot path (ri.rs:356-408) --
    def ao2mo(self, eigenvector: MatrixFull) -> "RIFull":
        return self.ao2mo_v02(eigenvector)

    def ao2mo_v02(self, eigenvector: MatrixFull) -> "RIFull":
        """AO(num_basis, num_basis, num_auxbas) -> MO(num_auxbas, num_states, num_states), ri.rs:382-408"""
        num_basis, num_states = eigenvector.size[0], eigenvector.size[1]
        num_auxbas = self.size[2]
        ri3mo = RIFull.new([num_auxbas, num_states, num_states], 0.0)
        ri_ao2mo_f(eigenvector.data, self.data, ri3mo.data, num_states, num_basis, num_auxbas)
        return ri3mo

    def ao2mo_v01(self, eigenvector: MatrixFull) -> "RIFull":
        """ri.rs:360-379 computes the same tensor with hand-rolled loops; here it is the same CUDA path."""
        return self.ao2mo_v02(eigenvector)

    def ao2mo_rect(self, c_left: MatrixFull, c_right: MatrixFull) -> "RIFull":
        """North-star occ-vir form: out[P,a,b] = sum C_L[mu,a] A[mu,nu,P] C_R[nu,b] (not in the reference)."""
        nb, nl, nr, nx = c_left.size[0], c_left.size[1], c_right.size[1], self.size[2]
        if c_right.size[0] != nb or self.size[0] != nb or self.size[1] != nb:
            raise ValueError("ao2mo_rect: inconsistent shapes")
        out = RIFull.new([nx, nl, nr], 0.0)
        check(lib.rb_host_ri_ao2mo(_ptr(c_left.data), nl, _ptr(c_right.data), nr, _ptr(self.data), _ptr(out.data), nb,
                                   nx), "ao2mo_rect")
        return out

    def ao2mo_jk(self, eigenvector: MatrixFull, dm: MatrixFull, ct: MatrixFull, ri3mo: Optional["RIFull"] = None):
        """One streaming pass over this (host) tensor: ao2mo + d_P + J + K with every P-chunk uploaded once
        (H2D | DMMA GEMMs | D2H overlapped).  Returns (ri3mo, d, J, K).  Not in the reference: REST makes the
        equivalent calls one by one (ao2mo, then _dgemv/_dgemm/_dsyrk per slab)."""
        nb, ns, nx, no = eigenvector.size[0], eigenvector.size[1], self.size[2], ct.size[1]
        if ri3mo is None:
            ri3mo = RIFull.new([nx, ns, ns], 0.0)
        d = np.zeros(nx, dtype=np.float64)
        j = MatrixFull.new([nb, nb], 0.0)
        k = MatrixFull.new([nb, nb], 0.0)
        check(lib.rb_host_ri_ao2mo_jk(_ptr(eigenvector.data), ns, _ptr(eigenvector.data), ns, _ptr(self.data),
                                      _ptr(ri3mo.data), nb, nx, _ptr(dm.data), _ptr(ct.data), no, _ptr(d), _ptr(j.data),
                                      _ptr(k.data)), "ao2mo_jk")
        return ri3mo, d, j, k

    def ao2mo_jk_upper(self, eigenvector: MatrixFull, dm: MatrixFull, ct: MatrixFull, symmetric_slabs: bool = False):
        """ao2mo_jk shipping only the a <= b part of ri3mo: returns (upper, d, J, K) with upper a float64 array [nx, ns(ns+1)/2],
        upper[P + nx * (b(b+1)/2 + a)] = ri3mo[P, a, b] (MatrixUpper's pair index per P).  Symmetric slabs make ri3mo symmetric in
        (a, b), so this is the whole tensor at half the device -> host traffic (rb_host_ri_ao2mo_jk_upper).  symmetric_slabs=True is
        the caller's guarantee that self[mu, nu, P] == self[nu, mu, P]: then only mu <= nu is uploaded as well
        (rb_host_ri_ao2mo_jk_symm; the strict lower triangles are never read)."""
        nb, ns, nx, no = eigenvector.size[0], eigenvector.size[1], self.size[2], ct.size[1]
        upper = np.zeros(nx * (ns * (ns + 1) // 2), dtype=np.float64)
        d = np.zeros(nx, dtype=np.float64)
        j = MatrixFull.new([nb, nb], 0.0)
        k = MatrixFull.new([nb, nb], 0.0)
        fn = lib.rb_host_ri_ao2mo_jk_symm if symmetric_slabs else lib.rb_host_ri_ao2mo_jk_upper
        check(fn(_ptr(eigenvector.data), ns, _ptr(self.data), _ptr(upper), nb, nx, _ptr(dm.data), _ptr(ct.data), no, _ptr(d),
                 _ptr(j.data), _ptr(k.data)), "ao2mo_jk_upper")
        return upper, d, j, k

    # -- slab copies (ri.rs:410-433) --
    def copy_from_ri(self, range_x: Range, range_y: Range, range_z: Range, from_ri: "RIFull", f_range_x: Range,
                     f_range_y: Range, f_range_z: Range) -> None:
        ri_copy_from_ri(from_ri.data, from_ri.size, f_range_x, f_range_y, f_range_z, self.data, self.size, range_x,
                        range_y, range_z)

    def copy_from_matr(self, range_x: Range, range_y: Range, i_z: int, copy_mod: int, from_matr: MatrixFull,
                       f_range_x: Range, f_range_y: Range) -> None:
        ri_copy_from_matr(from_matr.data, from_matr.size, f_range_x, f_range_y, self.data, self.size, range_x, range_y,
                          i_z, copy_mod)

    # -- RI-J / RI-K / d_P over this tensor (REST composes them from the primitives above) --
    def ri_dp(self, dm: MatrixFull) -> np.ndarray:
        nb, nx = self.size[0], self.size[2]
        d = np.zeros(nx, dtype=np.float64)
        check(lib.rb_host_ri_dp(_ptr(self.data), _ptr(dm.data), _ptr(d), nb, nx), "ri_dp")
        return d

    def ri_j(self, d: np.ndarray) -> MatrixFull:
        nb, nx = self.size[0], self.size[2]
        j = MatrixFull.new([nb, nb], 0.0)
        check(lib.rb_host_ri_j(_ptr(self.data), _ptr(_f64(d)), _ptr(j.data), nb, nx), "ri_j")
        return j

    def ri_k(self, ct: MatrixFull) -> MatrixFull:
        nb, nx, no = self.size[0], self.size[2], ct.size[1]
        k = MatrixFull.new([nb, nb], 0.0)
        check(lib.rb_host_ri_k(_ptr(self.data), _ptr(ct.data), no, _ptr(k.data), nb, nx), "ri_k")
        return k


    # -- (ia|jb)-type consumers of a P-fastest MO tensor --
    def ri_iajb(self, range_l_a: Range, range_r_a: Range, range_l_b: Range, range_r_b: Range,
                other: Optional["RIFull"] = None) -> MatrixFull:
        """self = ri3mo[P, l, r] (the output layout of ao2mo, src/ri.rs:381-386), other = a second such tensor (alpha /
        beta spin blocks; default self).  Returns the [|l_a||r_a|, |l_b||r_b|] matrix
        sum_P self[P,l,r] * other[P,l',r'] with row index (l - l0) + (r - r0)*|l| -- REST's (ia|jb) blocks."""
        b = self if other is None else other
        if b.size[0] != self.size[0]:
            raise RestB200Error("ri_iajb: the two MO tensors have different auxiliary dimensions")
        for t, (rl_, rr_) in ((self, (range_l_a, range_r_a)), (b, (range_l_b, range_r_b))):
            if not (0 <= rl_[0] <= rl_[1] <= t.size[1] and 0 <= rr_[0] <= rr_[1] <= t.size[2]):
                raise RestB200Error("ri_iajb: box outside the tensor")  # the Rust slicing would panic
        m, n = _rlen(range_l_a) * _rlen(range_r_a), _rlen(range_l_b) * _rlen(range_r_b)
        out = MatrixFull.new([m, n], 0.0)
        check(lib.rb_host_ri_iajb(self.size[0], _ptr(self.data), self.size[1], self.size[2], range_l_a[0],
                                  _rlen(range_l_a), range_r_a[0], _rlen(range_r_a), _ptr(b.data), b.size[1], b.size[2],
                                  range_l_b[0], _rlen(range_l_b), range_r_b[0], _rlen(range_r_b), _ptr(out.data)),
              "ri_iajb")
        return out

    def ri_mo_pq(self, range_l: Range, range_r: Range, w: Optional[np.ndarray] = None) -> MatrixFull:
        """self = ri3mo[P, l, r]; returns the symmetric [naux, naux] matrix sum_{(l,r) in box} w[l,r] self[P,l,r]
        self[Q,l,r] (w over the box's pairs, l fastest; None = ones) -- e.g. REST's RPA polarisability."""
        if not (0 <= range_l[0] <= range_l[1] <= self.size[1] and 0 <= range_r[0] <= range_r[1] <= self.size[2]):
            raise RestB200Error("ri_mo_pq: box outside the tensor")
        ll, rl = _rlen(range_l), _rlen(range_r)
        wv = None
        if w is not None:
            wv = _f64(w)
            if wv.size != ll * rl:
                raise RestB200Error("ri_mo_pq: one weight per MO pair of the box is needed")
        out = MatrixFull.new([self.size[0], self.size[0]], 0.0)
        check(lib.rb_host_ri_mo_pq(_ptr(self.data), self.size[0], self.size[1], self.size[2], range_l[0], ll, range_r[0],
                                   rl, None if wv is None else _ptr(wv), _ptr(out.data)), "ri_mo_pq")
        return out

# ======================================================================================================
# views and index maps
# ======================================================================================================
class MatrixFullSlice:
    """src/matrix/matrixfullslice.rs:184-188: borrowed column-major view (``size``, ``indicing``, ``data`` slices)."""

    def __init__(self, size, indicing, data: np.ndarray):
        self.size = [int(size[0]), int(size[1])]
        self.indicing = list(indicing)
        self.data = data

    def get_slice_x(self, y: int) -> np.ndarray:
        """matrixfullslice.rs:292-296"""
        return self.data[self.indicing[1] * y: self.indicing[1] * (y + 1)]

    def transpose(self) -> MatrixFull:
        """matrixfullslice.rs:298-318"""
        r, c = self.size
        out = MatrixFull.new([c, r], 0.0)
        if r * c:
            src = np.ascontiguousarray(self.data[: r * c])
            check(lib.rb_host_matrix_transpose(_ptr(src), r, c, _ptr(out.data)), "MatrixFullSlice::transpose")
        return out

    transpose_and_drop = transpose

    def ddot(self, b: "MatrixFullSlice") -> Optional[MatrixFull]:
        """matrix_blas_lapack.rs:714-730: a*b accumulated (beta = 1) onto a zeroed C; None on a shape mismatch"""
        if self.size[1] != b.size[0]:
            return None
        m, n, k = self.size[0], b.size[1], self.size[1]
        c = MatrixFull.new([m, n], 0.0)
        check(lib.rb_host_dgemm(b'N', b'N', m, n, k, 1.0, _ptr(self.data), max(m, 1), _ptr(b.data), max(k, 1), 1.0,
                                _ptr(c.data), max(m, 1)), "MatrixFullSlice::ddot")
        return c


class MatrixFullSliceMut(MatrixFullSlice):
    """src/matrix/matrixfullslice.rs (MatrixFullSliceMut): the mutable view; writes land in the parent's ``data``."""

    def lapack_dgemm(self, a: MatrixFullSlice, b: MatrixFullSlice, opa: str, opb: str, alpha: float, beta: float) -> None:
        """matrix_blas_lapack.rs:739-774: c = alpha*op(a)*op(b) + beta*c (the reference's shape check is disabled)"""
        m = a.size[0] if opa == 'N' else a.size[1]
        k = a.size[1] if opa == 'N' else a.size[0]
        n = b.size[1] if opb == 'N' else b.size[0]
        lda = max(m, 1) if opa == 'N' else max(k, 1)
        ldb = max(k, 1) if opb == 'N' else max(n, 1)
        check(lib.rb_host_dgemm(ch(opa), ch(opb), m, n, k, alpha, _ptr(a.data), lda, _ptr(b.data), ldb, beta,
                                _ptr(self.data), max(m, 1)), "MatrixFullSliceMut::lapack_dgemm")


class MatrixUpperSlice:
    """src/matrix/matrixupper.rs:507-520: borrowed packed-upper view."""

    def __init__(self, data: np.ndarray):
        self.size = int(data.size)
        self.data = data

    @staticmethod
    def from_vec(new_vec: np.ndarray) -> "MatrixUpperSlice":
        return MatrixUpperSlice(new_vec)

    def to_matrixfull(self) -> Optional[MatrixFull]:
        """matrixupper.rs:521-561: unpack + mirror; None unless the length is triangular"""
        n = int((1.0 + 8.0 * float(self.size)) ** 0.5 * 0.5 - 0.5)
        if n * (n + 1) // 2 != self.size:
            return None
        if self.size == 0:
            return MatrixFull.empty()
        out = MatrixFull.new([n, n], 0.0)
        src = np.ascontiguousarray(self.data)
        check(lib.rb_host_to_matrixfull(_ptr(src), self.size, _ptr(out.data)), "MatrixUpperSlice::to_matrixfull")
        return out


class ERIFold4:
    """src/eri.rs:170-373: (ij|kl) with both index pairs folded -- a column-major [npair_ij, npair_kl] tensor,
    ``indicing = [1, size[0]]``, packed pair index ``j(j+1)/2 + i``.  The chunk scatters run on the GPU (csrc/rb_eri.cu)."""

    def __init__(self, size, data: np.ndarray):
        self.size = [int(size[0]), int(size[1])]
        self.indicing = [1, self.size[0]]
        self.data = data

    @staticmethod
    def new(size, new_default: float) -> "ERIFold4":
        return ERIFold4(size, np.full(int(size[0]) * int(size[1]), float(new_default), dtype=np.float64))

    @staticmethod
    def from_vec_unchecked(size, new_vec) -> "ERIFold4":
        return ERIFold4(size, _f64(new_vec))

    @staticmethod
    def from_vec(size, new_vec) -> Optional["ERIFold4"]:
        """eri.rs:204-218: panics when the vector is shorter than the tensor, warns (and keeps it) when it is longer"""
        t = ERIFold4.from_vec_unchecked(size, new_vec)
        ln = t.indicing[1] * t.size[1]
        if ln > t.data.size:
            raise ValueError("Error: inconsistency happens when formating a tensor from a given vector, (length from size, "
                             f"length of new vector) = ({ln},{t.data.size})")
        if ln < t.data.size:
            print(f"Waring: the vector size ({t.data.size}) is larger for the size of the new tensor ({ln})")
        return t

    def get_reducing_matrix(self, i_reduced: int) -> "MatrixUpperSlice":
        """eri.rs:229-236 (and the `_mut` form 219-227): column i_reduced as a packed-upper view (zero copy)"""
        if not 0 <= i_reduced < self.size[1]:
            raise IndexError("ERIFold4::get_reducing_matrix: index2d([0, i_reduced]) is None")
        p0 = i_reduced * self.indicing[1]
        v = MatrixUpperSlice(self.data[p0:p0 + self.indicing[1]])
        v.size = self.size[0]
        return v

    get_reducing_matrix_mut = get_reducing_matrix

    def _scatter(self, ranges, buf, mode: int, what: str) -> None:
        (i0, i1), (j0, j1), (k0, k1), (l0, l1) = [(int(r[0]), int(r[1])) for r in ranges]
        need = (i1 - i0) * (j1 - j0) * (k1 - k0) * (l1 - l0)
        b = _f64(buf)
        if b.size < need:
            raise ValueError(f"{what}: the local block holds {b.size} elements, the ranges describe {need}")
        check(lib.rb_host_erifold4_chunk_copy(_ptr(self.data), self.size[0], self.size[1], i0, i1 - i0, j0, j1 - j0, k0, k1 - k0,
                                              l0, l1 - l0, _ptr(b), mode), what)

    def chunk_copy_from_local_erifull(self, dim: int, d1: Range, d2: Range, d3: Range, d4: Range, buf) -> None:
        """eri.rs:266-305: dense local block [|d1|, |d2|, |d3|, |d4|] -> the elements with k <= l and i <= j.  `dim` is the number
        of basis functions: dim(dim+1)/2 must be the column length (the reference computes slice starts with it)."""
        if dim * (dim + 1) // 2 != self.size[0]:
            raise ValueError("chunk_copy_from_local_erifull: dim(dim+1)/2 differs from the column length of the tensor")
        self._scatter((d1, d2, d3, d4), buf, 0, "ERIFold4::chunk_copy_from_local_erifull")

    def chunk_copy_from_a_full_vector(self, ranges, buf) -> None:
        """eri.rs:308-372: the libcint shell-quartet scatter (ranges[0].start < ranges[1].start: whole (i, j) block;
        equal starts: the local upper triangle; otherwise nothing)"""
        self._scatter(ranges, buf, 1, "ERIFold4::chunk_copy_from_a_full_vector")


class MatrixUpperStepBy:
    """src/matrix/matrix_trait.rs:170-217: iterator adaptor that walks a column-major n x n sequence and keeps the
    positions with row <= column (the packed-upper order).  ``iter`` yields linear positions; with ``source`` the items
    are the elements at those positions (the `Iter<T>` form), without it the positions themselves (`IterMut` form)."""

    def __init__(self, it, size, shift: int = 0, source: Optional[np.ndarray] = None):
        self.iter = it
        self.size = [int(size[0]), int(size[1])]
        self.step = self.size[0]
        self.position = int(shift)
        self.first_take = True
        self._source = source

    @staticmethod
    def new_shift(it, size, shift: int) -> "MatrixUpperStepBy":
        return MatrixUpperStepBy(it, size, shift)

    def _nth(self, n: int):
        v = None
        for _ in range(n + 1):
            v = next(self.iter)  # StopIteration ends the walk exactly like `None` from `nth`
        return v

    def __iter__(self):
        return self

    def __next__(self):
        curr_row = self.position % self.size[0]
        curr_column = self.position // self.size[0]
        if self.first_take:
            self.position += 1
            self.first_take = False
            pos = self._nth(self.position - 1)
        elif curr_row <= curr_column:
            self.position += 1
            pos = next(self.iter)
        else:
            step = self.size[0] - curr_column
            self.position += step
            pos = self._nth(step - 1)
        return pos if self._source is None else float(self._source[pos])


def map_upper_to_full(size: int) -> Optional[np.ndarray]:
    """matrixupper.rs:587-604: packed index -> [i, j] (an int array [size, 2], the `MatrixUpper<[usize;2]>` data);
    None unless ``size`` is triangular.  Integer index arithmetic only."""
    n = int((1.0 + 8.0 * float(size)) ** 0.5 * 0.5 - 0.5)
    if n * (n + 1) // 2 != size:
        return None
    j = np.repeat(np.arange(n, dtype=np.int64), np.arange(1, n + 1, dtype=np.int64))
    i = np.arange(size, dtype=np.int64) - j * (j + 1) // 2
    return np.stack([i, j], axis=1)


def map_full_to_upper(size) -> Optional[np.ndarray]:
    """matrixupper.rs:605-617: n x n column-major int matrix holding the packed index at every i <= j (0 elsewhere);
    None unless square.  (The reference fills it through iter_matrixupper_mut, which is None for n == 0.)"""
    if size[0] != size[1]:
        return None
    n = int(size[0])
    out = np.zeros(n * n, dtype=np.int64)
    if n:
        m = map_upper_to_full(n * (n + 1) // 2)
        out[m[:, 0] + m[:, 1] * n] = np.arange(m.shape[0], dtype=np.int64)
    return out


# ======================================================================================================
# matrix_blas_lapack (src/matrix/matrix_blas_lapack.rs)
# ======================================================================================================
def general_check_shape(matr_a, matr_b, opa: str, opb: str) -> bool:
    """matrix_blas_lapack.rs:14-34 (as written there: equal sizes for NN / TT, reversed-equal for TN / NT)"""
    sa, sb = list(matr_a.size), list(matr_b.size)
    if (opa, opb) in (('N', 'N'), ('T', 'T')):
        return all(a == b for a, b in zip(sa, sb))
    if (opa, opb) in (('T', 'N'), ('N', 'T')):
        return all(a == b for a, b in zip(reversed(sa), sb))
    return False


def _dgemm_nn(mat_a, mat_b) -> MatrixFull:
    """matrix_blas_lapack.rs:1185-1203: c = a*b.  The reference's rayon / scalar loops (and `_dgemm_nn_serial`,
    1206-1221) define only a summation order; here all of them are the DMMA GEMM (1e-10 parity, not bitwise)."""
    (ax, ay), (bx, by) = mat_a.size, mat_b.size
    if ay != bx:
        raise ValueError("For the input matrices: mat_a[ax,ay], mat_b[bx,by], ay!=bx. dgemm false")
    c = MatrixFull.new([ax, by], 0.0)
    if ax == 0 or by == 0:
        return c
    check(lib.rb_host_dgemm(b'N', b'N', ax, by, ay, 1.0, _ptr(mat_a.data), max(ax, 1), _ptr(mat_b.data), max(bx, 1), 0.0,
                            _ptr(c.data), max(ax, 1)), "_dgemm_nn")
    return c


_dgemm_nn_serial = _dgemm_nn


def _dgemm_tn(mat_a, mat_b) -> MatrixFull:
    """matrix_blas_lapack.rs:1224-1237 (`_dgemm_tn_serial`: 1240-1253): c = a^T*b"""
    (ax, ay), (bx, by) = mat_a.size, mat_b.size
    if ax != bx:
        raise ValueError("For the input matrices: mat_a[ax,ay], mat_b[bx,by], ay!=bx. dgemm false")
    c = MatrixFull.new([ay, by], 0.0)
    if ay == 0 or by == 0:
        return c
    check(lib.rb_host_dgemm(b'T', b'N', ay, by, ax, 1.0, _ptr(mat_a.data), max(ax, 1), _ptr(mat_b.data), max(bx, 1), 0.0,
                            _ptr(c.data), max(ay, 1)), "_dgemm_tn")
    return c


_dgemm_tn_serial = _dgemm_tn


def _dgemm_tn_v02(mat_a, mat_b, to_slice) -> None:
    """matrix_blas_lapack.rs:1254-1271: c = a^T*b written through ``to_slice`` -- the flattened x-runs of
    RIFull::get_slices_mut (a list of writable views), column-major [ay, by] across the concatenation."""
    (ax, ay), (bx, by) = mat_a.size, mat_b.size
    if ax != bx:
        raise ValueError("For the input matrices: mat_a[ax,ay], mat_b[bx,by], ay!=bx. dgemm false")
    c = _dgemm_tn(mat_a, mat_b)
    off = 0
    for run in to_slice:  # zip semantics: stops at the shorter of (destination, ay*by)
        n = min(run.size, c.data.size - off)
        if n <= 0:
            break
        run[:n] = c.data[off:off + n]
        off += n


# -- einsum helpers (matrix_blas_lapack.rs:1273-1387, matrix/einsum.rs) --
def _einsum_01_rayon(mat_a, vec_b: np.ndarray) -> MatrixFull:
    """matrix_blas_lapack.rs:1275-1290: "ij,j->ij"; the first len(vec_b) columns of mat_a scaled by vec_b"""
    i_len, j_len = mat_a.size[0], int(np.asarray(vec_b).size)
    om = MatrixFull.new([i_len, j_len], 0.0)
    if i_len == 0 or j_len == 0:
        return om
    if j_len > mat_a.size[1]:
        raise ValueError("einsum ij,j->ij: vec_b longer than the number of columns of mat_a")  # `.unwrap()` on None
    b = _f64(vec_b)
    check(lib.rb_host_einsum(1, _ptr(mat_a.data), _ptr(b), _ptr(om.data), i_len, j_len), "_einsum_01")
    return om


_einsum_01_serial = _einsum_01_rayon


def _einsum_02_rayon(mat_a, mat_b) -> np.ndarray:
    """matrix_blas_lapack.rs:1293-1311: "ip,ip->p" over the common columns (zip stops at the shorter operand)"""
    (a_x, a_y), (b_x, b_y) = mat_a.size, mat_b.size
    n_p = min(a_y, b_y)
    out = np.zeros(n_p, dtype=np.float64)
    if a_x == 0 or b_x == 0 or n_p == 0:
        return out
    if a_x != b_x:  # zip of unequal columns truncates to the shorter one in the reference; not a hot-path shape
        raise ValueError("einsum ip,ip->p: operands with different row counts are not supported on the device path")
    check(lib.rb_host_einsum(2, _ptr(mat_a.data), _ptr(mat_b.data), _ptr(out), a_x, n_p), "_einsum_02")
    return out


_einsum_02_serial = _einsum_02_rayon


def _einsum_03(vec_a: np.ndarray, vec_b: np.ndarray) -> MatrixFull:
    """matrix_blas_lapack.rs:1355-1387 (`_einsum_03`, `_einsum_03_forvec`): "i,j->ij" outer product"""
    a, b = _f64(vec_a), _f64(vec_b)
    om = MatrixFull.new([a.size, b.size], 0.0)
    if a.size and b.size:
        check(lib.rb_host_einsum(3, _ptr(a), _ptr(b), _ptr(om.data), a.size, b.size), "_einsum_03")
    return om


_einsum_03_forvec = _einsum_03


def _einsum_general(mat_a, mat_b, opt: str) -> MatrixFull:
    """matrix/einsum.rs:3-14: dispatch on the subscript string; panics on anything else"""
    if opt == "ij,j->ij":
        return _einsum_01_rayon(mat_a, mat_b.data)
    if opt == "ip,ip->p":
        v = _einsum_02_rayon(mat_a, mat_b)
        return MatrixFull.from_vec([v.size, 1], v)
    if opt == "i,j->ij":
        return _einsum_03(mat_a.data, mat_b.data)
    if opt == "ij,ji->ij":  # einsum.rs:81-91 (`_einsum_04_general`): despite the label it is a plain dgemm NN
        c = MatrixFull.new([mat_a.size[0], mat_b.size[1]], 0.0)
        _dgemm_full(mat_a, 'N', mat_b, 'N', c, 1.0, 0.0)
        return c
    raise ValueError(f"Not implemented for einsum: {opt}")


def _gemm_shape_ok(sa, opa, sb, opb, sc) -> bool:
    key = (opa, opb)
    if key == ('N', 'N'):
        return sa[1] == sb[0] and sa[0] == sc[0] and sb[1] == sc[1]
    if key == ('T', 'N'):
        return sa[0] == sb[0] and sa[1] == sc[0] and sb[1] == sc[1]
    if key == ('N', 'T'):
        return sa[1] == sb[1] and sa[0] == sc[0] and sb[0] == sc[1]
    if key == ('T', 'T'):
        return sa[0] == sb[1] and sa[1] == sc[0] and sb[0] == sc[1]
    return False


def _dgemm(matr_a: MatrixFull, sub_a_dim, opa: str, matr_b: MatrixFull, sub_b_dim, opb: str, matr_c: MatrixFull,
           sub_c_dim, alpha: float, beta: float) -> None:
    """matrix_blas_lapack.rs:122-178: sub-block GEMM; same three panics (shape, within, contiguity)."""
    la = [_rlen(sub_a_dim[0]), _rlen(sub_a_dim[1])]
    lb = [_rlen(sub_b_dim[0]), _rlen(sub_b_dim[1])]
    lc = [_rlen(sub_c_dim[0]), _rlen(sub_c_dim[1])]
    if not _gemm_shape_ok(la, opa, lb, opb, lc):
        raise ValueError(f"ERROR:: Matr_A[{la[0]},{la[1]},{opa}] * Matr_B[{lb[0]},{lb[1]},{opb}] -> Matr_C[{lc[0]},{lc[1]}]")
    within = (sub_a_dim[0][1] <= matr_a.size[0] and sub_a_dim[1][1] <= matr_a.size[1]
              and sub_b_dim[0][1] <= matr_b.size[0] and sub_b_dim[1][1] <= matr_b.size[1]
              and sub_c_dim[0][1] <= matr_c.size[0] and sub_c_dim[1][1] <= matr_c.size[1])
    if not within:
        raise ValueError("ERROR:: sub-matrix block is not within the matrix")
    general_dgemm_f(matr_a.data, matr_a.size, sub_a_dim[0], sub_a_dim[1], opa,
                    matr_b.data, matr_b.size, sub_b_dim[0], sub_b_dim[1], opb,
                    matr_c.data, matr_c.size, sub_c_dim[0], sub_c_dim[1], alpha, beta)


def _dgemm_full(matr_a: MatrixFull, opa: str, matr_b: MatrixFull, opb: str, matr_c: MatrixFull, alpha: float,
                beta: float) -> None:
    """matrix_blas_lapack.rs:180-252"""
    sa, sb, sc = matr_a.size, matr_b.size, matr_c.size
    if not _gemm_shape_ok(sa, opa, sb, opb, sc):
        raise ValueError(f"ERROR:: Matr_A[{sa[0]},{sa[1]},{opa}] * Matr_B[{sb[0]},{sb[1]},{opb}] -> Matr_C[{sc[0]},{sc[1]}]")
    m = sa[0] if opa == 'N' else sa[1]
    k = sa[1] if opa == 'N' else sa[0]
    n = sb[1] if opb == 'N' else sb[0]
    lda = max(m, 1) if opa == 'N' else max(k, 1)
    ldb = max(k, 1) if opb == 'N' else max(n, 1)
    check(lib.rb_host_dgemm(ch(opa), ch(opb), m, n, k, alpha, _ptr(matr_a.data), lda, _ptr(matr_b.data), ldb, beta,
                            _ptr(matr_c.data), max(m, 1)), "_dgemm_full")


def _dgemm_full_new(matr_a: MatrixFull, opa: str, matr_b: MatrixFull, opb: str, alpha: float, beta: float) -> MatrixFull:
    """matrix_blas_lapack.rs:256-278"""
    axy, bxy = matr_a.size, matr_b.size
    if axy[0] == 0 or bxy[1] == 0:
        return MatrixFull.new([axy[0], bxy[1]], 0.0)
    shape = {('N', 'N'): [axy[0], bxy[1]], ('T', 'N'): [axy[1], bxy[1]], ('N', 'T'): [axy[0], bxy[0]],
             ('T', 'T'): [axy[1], bxy[0]]}.get((opa, opb), [0, 0])
    c = MatrixFull.new(shape, 0.0)
    _dgemm_full(matr_a, opa, matr_b, opb, c, alpha, beta)
    return c


def _dsyrk(matr_a: MatrixFull, matr_c: MatrixFull, uplo: str, trans: str, alpha: float, beta: float) -> None:
    """matrix_blas_lapack.rs:392-413: panics unless C square; only the uplo triangle of C is read/written"""
    m, n = matr_c.size
    if m != n:
        raise ValueError("matr_b should be symmetric")
    is_n = trans.lower() == 'n'
    k = matr_a.size[1] if is_n else matr_a.size[0]
    lda = max(n, 1) if is_n else max(k, 1)
    check(lib.rb_host_dsyrk(ch(uplo), ch(trans), n, k, alpha, _ptr(matr_a.data), lda, beta, _ptr(matr_c.data),
                            max(n, 1)), "_dsyrk")


def _dsymm(matr_a: MatrixFull, matr_b: MatrixFull, matr_c: MatrixFull, side: str, uplo: str, alpha: float,
           beta: float) -> None:
    """matrix_blas_lapack.rs:354-378 (no checks, like the reference)"""
    m, n = matr_c.size
    lda = m if side.lower() == 'l' else n
    check(lib.rb_host_dsymm(ch(side), ch(uplo), m, n, alpha, _ptr(matr_a.data), lda, _ptr(matr_b.data), m, beta,
                            _ptr(matr_c.data), m), "_dsymm")


def _dgemv(matr_a: MatrixFull, vec_x: np.ndarray, vec_y: np.ndarray, trans: str, alpha: float, beta: float, incx: int,
           incy: int) -> None:
    """matrix_blas_lapack.rs:38-70: same length checks (=> panic) as the reference"""
    m, n = matr_a.size
    is_n = trans.lower() == 'n'
    ok_x = vec_x.size == 1 + ((n if is_n else m) - 1) * abs(incx)
    ok_y = vec_y.size == 1 + ((m if is_n else n) - 1) * abs(incy)
    if not (ok_x and ok_y):
        raise ValueError(f"ERROR:: Matr_A[{m},{n},{trans}] * Vec_X[{vec_x.size}] -> Vec_Y[{vec_y.size}]")
    check(lib.rb_host_dgemv(ch(trans), m, n, alpha, _ptr(matr_a.data), max(m, 1), _ptr(vec_x), incx, beta, _ptr(vec_y),
                            incy), "_dgemv")


# ======================================================================================================
# eigen-solvers (matrix_blas_lapack.rs:319-352, 599-652, 2123-2185)
# ======================================================================================================
def _dsyev(matr_a: MatrixFull, jobz: str):
    """(Some(eigenvectors) | None, eigenvalues ascending, n); panics (raises) for a non-square matrix like the reference"""
    n = matr_a.size[0]
    if matr_a.size[0] != matr_a.size[1]:
        raise RestB200Error("Error in _dsyev: the algorithm is only vaild for real symmetric matrices")
    w = np.zeros(n, dtype=np.float64)
    vec = MatrixFull.new([n, n], 0.0) if jobz in ("V", "v") else None
    check(lib.rb_host_dsyev(ch(jobz), n, _ptr(matr_a.data), _ptr(w), None if vec is None else _ptr(vec.data)), "_dsyev")
    return vec, w, n


def _power(matr_a: MatrixFull, p: float, threshold: float) -> Optional[MatrixFull]:
    """A^p over the eigenvalues >= threshold (e.g. p = -0.5: S^-1/2); None for a non-square matrix"""
    if matr_a.size[0] != matr_a.size[1]:
        return None
    n = matr_a.size[0]
    om = MatrixFull.new([n, n], 0.0)
    kept = C.c_int(0)
    check(lib.rb_host_power(n, _ptr(matr_a.data), float(p), float(threshold), _ptr(om.data), C.byref(kept)), "_power")
    return om


def _dspgvx(matr_a: MatrixUpper, matr_b: MatrixUpper, num_orb: int):
    if matr_a.data.size != matr_b.data.size:
        raise RestB200Error("ERROR:: _dspgvx for BasicMatUp, Matr_A and Matr_B have different size")
    n = matr_a.size2d()[0]
    if not 0 <= num_orb <= n:
        raise RestB200Error("Error:: The number of outcoming eigenvectors is unequal to the orbital number")
    w = np.zeros(num_orb, dtype=np.float64)
    z = MatrixFull.new([n, num_orb], 0.0)
    check(lib.rb_host_dspgvx(n, _ptr(matr_a.data), _ptr(matr_b.data), num_orb, _ptr(w), _ptr(z.data)), "_dspgvx")
    return z, w
