"""rest_tensors_b200 -- B200 (sm_100a) implementation of the rest_tensors RI hot path.

Layout of the package (only what the hot path needs):
  csrc/         hand-written CUDA kernels + the C ABI (include/rest_b200.h) -> librest_b200.so
  _lib.py       ctypes binding of the C ABI (import fails loudly if the .so is missing)
  tensors.py    host-side mirror of the reference API (RIFull / MatrixFull / MatrixUpper, _dgemm*, _dsyrk, ...)
  device.py     device-resident API on torch-owned HBM buffers + the P-sharded RI tensor
"""
from ._lib import lib, RestB200Error, LIB_PATH, SIGNATURES  # noqa: F401
from .tensors import (  # noqa: F401
    RIFull, MatrixFull, MatrixUpper, MatrixFullSlice, MatrixFullSliceMut, MatrixUpperSlice, MatrixUpperStepBy, ERIFold4,
    map_upper_to_full, map_full_to_upper, general_check_shape,
    _dgemm, _dgemm_full, _dgemm_full_new, _dsyrk, _dsymm, _dgemv,
    _dgemm_nn, _dgemm_nn_serial, _dgemm_tn, _dgemm_tn_serial, _dgemm_tn_v02,
    _einsum_01_rayon, _einsum_01_serial, _einsum_02_rayon, _einsum_02_serial, _einsum_03, _einsum_03_forvec, _einsum_general,
    ri_ao2mo_f, general_dgemm_f, special_dgemm_f_01, matr_copy, matr_copy_from_ri, ri_copy_from_matr, ri_copy_from_ri,
    _dsyev, _power, _dspgvx,
)

__all__ = [
    "RIFull", "MatrixFull", "MatrixUpper", "MatrixFullSlice", "MatrixFullSliceMut", "MatrixUpperSlice", "MatrixUpperStepBy", "ERIFold4",
    "map_upper_to_full", "map_full_to_upper", "general_check_shape", "RestB200Error",
    "_dgemm", "_dgemm_full", "_dgemm_full_new", "_dsyrk", "_dsymm", "_dgemv",
    "_dgemm_nn", "_dgemm_nn_serial", "_dgemm_tn", "_dgemm_tn_serial", "_dgemm_tn_v02",
    "_einsum_01_rayon", "_einsum_01_serial", "_einsum_02_rayon", "_einsum_02_serial", "_einsum_03", "_einsum_03_forvec",
    "_einsum_general",
    "ri_ao2mo_f", "general_dgemm_f", "special_dgemm_f_01", "matr_copy", "matr_copy_from_ri", "ri_copy_from_matr",
    "ri_copy_from_ri", "_dsyev", "_power", "_dspgvx",
]
