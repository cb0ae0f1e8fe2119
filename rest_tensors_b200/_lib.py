"""ctypes binding of librest_b200.so (the C ABI declared in include/rest_b200.h).

The library is the product: there is no Python/NumPy/PyTorch implementation of any operation in this
package.  If the shared object is missing the import fails loudly; if no CUDA device is present every
compute call fails with ``RestB200Error`` (there is no CPU fallback).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# REST_B200_LIB: development override used by tools/ to A/B kernel variants (still a build of csrc/)
LIB_PATH = os.environ.get("REST_B200_LIB") or os.path.join(_HERE, "librest_b200.so")


class RestB200Error(RuntimeError):
    """Raised when a librest_b200 call returns a non-zero status (the Rust wrappers panic on the same conditions)."""


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(or `make -C rest_tensors_b200/csrc`). rest_tensors_b200 has no CPU fallback."
    )

lib = C.CDLL(LIB_PATH)

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
c_i64 = C.c_int64
c_vp = C.c_void_p

# name -> (restype, argtypes); kept in one table so tests can check it against the header
SIGNATURES = {
    "rb_version": (C.c_int, []),
    "rb_last_error": (C.c_char_p, []),
    "rb_device_count": (C.c_int, []),
    "rb_ctx_create": (C.c_int, [C.c_int, C.POINTER(c_vp)]),
    "rb_ctx_destroy": (C.c_int, [c_vp]),
    "rb_ctx_set_stream": (C.c_int, [c_vp, c_vp]),
    "rb_ctx_use_own_stream": (C.c_int, [c_vp]),
    "rb_ctx_sync": (C.c_int, [c_vp]),
    "rb_ctx_num_sms": (C.c_int, [c_vp]),
    "rb_ctx_launch_count": (c_i64, [c_vp]),
    "rb_ctx_tma_layout_count": (c_i64, [c_vp]),
    "rb_ctx_set_layout_path": (C.c_int, [c_vp, C.c_int]),
    "rb_host_register": (C.c_int, [c_vp, c_i64]),
    "rb_host_unregister": (C.c_int, [c_vp]),
    "rb_gemm_plan_stream_k": (C.c_int, [c_i64, c_i64, c_i64, c_i64, C.c_int, C.c_int, c_vp]),
    "rb_gemm_stream_k_tables": (C.c_int, [c_i64, c_i64, c_i64, c_i64, C.c_int, C.c_int, c_vp, c_vp, c_vp, c_vp]),
    "rb_erifold4_chunk_copy": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64] + [C.c_int] * 8 + [c_vp, C.c_int]),
    "rb_host_erifold4_chunk_copy": (C.c_int, [c_vp, c_i64, c_i64] + [C.c_int] * 8 + [c_vp, C.c_int]),
    "rb_ctx_set_gemm_path": (C.c_int, [c_vp, C.c_int]),
    "rb_ctx_poison_workspaces": (C.c_int, [c_vp]),
    "rb_graph_begin": (C.c_int, [c_vp]),
    "rb_graph_end": (C.c_int, [c_vp, C.POINTER(c_vp)]),
    "rb_graph_launch": (C.c_int, [c_vp, c_vp]),
    "rb_graph_kernel_count": (C.c_int64, [c_vp]),
    "rb_graph_free": (C.c_int, [c_vp, c_vp]),
    "rb_dev_alloc": (C.c_int, [c_vp, c_i64, C.POINTER(c_vp)]),
    "rb_dev_free": (C.c_int, [c_vp, c_vp]),
    "rb_host_alloc_pinned": (C.c_int, [c_i64, C.POINTER(c_vp)]),
    "rb_host_free_pinned": (C.c_int, [c_vp]),
    "rb_peer_enable": (C.c_int, [c_vp, C.c_int]),
    "rb_ipc_export": (C.c_int, [c_vp, c_vp, c_vp, C.POINTER(c_i64)]),
    "rb_ipc_open": (C.c_int, [c_vp, c_vp, C.POINTER(c_vp)]),
    "rb_ipc_close": (C.c_int, [c_vp, c_vp]),
    "rb_memcpy_h2d": (C.c_int, [c_vp, c_vp, c_vp, c_i64]),
    "rb_memcpy_d2h": (C.c_int, [c_vp, c_vp, c_vp, c_i64]),
    # compat (Fortran ABI, everything by pointer)
    "ri_ao2mo_f_": (None, [c_vp, c_vp, c_vp, c_ip, c_ip, c_ip]),
    "general_dgemm_f_": (None, [c_vp, c_ip, c_ip, c_ip, c_ip, c_ip, c_ip, C.c_char_p,
                                 c_vp, c_ip, c_ip, c_ip, c_ip, c_ip, c_ip, C.c_char_p,
                                 c_vp, c_ip, c_ip, c_ip, c_ip, c_ip, c_ip, c_dp, c_dp]),
    "special_dgemm_f_01_": (None, [c_vp, c_ip, c_ip, c_ip, c_ip, c_ip, c_ip, c_ip, c_ip,
                                    c_vp, c_ip, c_ip, c_ip, c_ip, c_ip, c_ip, c_dp, c_dp]),
    "copy_mm_": (None, [c_ip, c_ip, c_vp, c_ip, c_ip, c_ip, c_ip, c_vp, c_ip, c_ip, c_ip, c_ip]),
    "copy_mr_": (None, [c_ip, c_ip, c_vp, c_ip, c_ip, c_ip, c_ip, c_vp, c_ip, c_ip, c_ip, c_ip, c_ip, c_ip, c_ip]),
    "copy_rm_": (None, [c_ip, c_ip, c_vp, c_ip, c_ip, c_ip, c_ip, c_ip, c_ip, c_ip, c_vp, c_ip, c_ip, c_ip, c_ip]),
    "copy_rr_": (None, [c_ip, c_ip, c_ip, c_vp, c_ip, c_ip, c_ip, c_ip, c_ip, c_ip,
                         c_vp, c_ip, c_ip, c_ip, c_ip, c_ip, c_ip]),
    # host-pointer wrappers
    "rb_host_dgemm": (C.c_int, [C.c_char, C.c_char, C.c_int, C.c_int, C.c_int, C.c_double, c_vp, C.c_int,
                                c_vp, C.c_int, C.c_double, c_vp, C.c_int]),
    "rb_host_dsyrk": (C.c_int, [C.c_char, C.c_char, C.c_int, C.c_int, C.c_double, c_vp, C.c_int, C.c_double,
                                c_vp, C.c_int]),
    "rb_host_dgemv": (C.c_int, [C.c_char, C.c_int, C.c_int, C.c_double, c_vp, C.c_int, c_vp, C.c_int,
                                C.c_double, c_vp, C.c_int]),
    "rb_host_dsymm": (C.c_int, [C.c_char, C.c_char, C.c_int, C.c_int, C.c_double, c_vp, C.c_int, c_vp, C.c_int,
                                C.c_double, c_vp, C.c_int]),
    "rb_host_to_matrixupper": (C.c_int, [c_vp, c_i64, c_vp]),
    "rb_host_to_matrixfull": (C.c_int, [c_vp, c_i64, c_vp]),
    "rb_host_ri_pack_symm": (C.c_int, [c_vp, c_i64, c_i64, c_vp]),
    "rb_host_ri_transpose": (C.c_int, [c_vp, c_i64, c_i64, c_i64, C.c_int, c_vp]),
    "rb_host_matrix_transpose": (C.c_int, [c_vp, c_i64, c_i64, c_vp]),
    "rb_host_ri_ao2mo": (C.c_int, [c_vp, C.c_int, c_vp, C.c_int, c_vp, c_vp, C.c_int, C.c_int]),
    "rb_host_ri_ao2mo_jk": (C.c_int, [c_vp, C.c_int, c_vp, C.c_int, c_vp, c_vp, C.c_int, C.c_int, c_vp, c_vp, C.c_int,
                                      c_vp, c_vp, c_vp]),
    "rb_host_ri_ao2mo_jk_upper": (C.c_int, [c_vp, C.c_int, c_vp, c_vp, C.c_int, C.c_int, c_vp, c_vp, C.c_int, c_vp, c_vp, c_vp]),
    "rb_host_ri_ao2mo_jk_symm": (C.c_int, [c_vp, C.c_int, c_vp, c_vp, C.c_int, C.c_int, c_vp, c_vp, C.c_int, c_vp, c_vp, c_vp]),
    "rb_host_axpy": (C.c_int, [C.c_int, c_vp, c_vp, C.c_double, C.c_double, c_i64]),
    "rb_host_ri_dp": (C.c_int, [c_vp, c_vp, c_vp, C.c_int, C.c_int]),
    "rb_host_ri_j": (C.c_int, [c_vp, c_vp, c_vp, C.c_int, C.c_int]),
    "rb_host_ri_k": (C.c_int, [c_vp, c_vp, C.c_int, c_vp, C.c_int, C.c_int]),
    # device-pointer API
    "rb_dgemm": (C.c_int, [c_vp, C.c_char, C.c_char, C.c_int, C.c_int, C.c_int, C.c_double, c_vp, c_i64,
                           c_vp, c_i64, C.c_double, c_vp, c_i64]),
    "rb_dgemm_strided_batched": (C.c_int, [c_vp, C.c_char, C.c_char, C.c_int, C.c_int, C.c_int, C.c_double,
                                           c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, C.c_double, c_vp, c_i64, c_i64,
                                           C.c_int]),
    "rb_dsyrk": (C.c_int, [c_vp, C.c_char, C.c_char, C.c_int, C.c_int, C.c_double, c_vp, c_i64, C.c_double,
                           c_vp, c_i64]),
    "rb_dgemv": (C.c_int, [c_vp, C.c_char, C.c_int, C.c_int, C.c_double, c_vp, c_i64, c_vp, C.c_int, C.c_double,
                           c_vp, C.c_int]),
    "rb_dsymm": (C.c_int, [c_vp, C.c_char, C.c_char, C.c_int, C.c_int, C.c_double, c_vp, c_i64, c_vp, c_i64,
                           C.c_double, c_vp, c_i64]),
    "rb_ri_ao2mo": (C.c_int, [c_vp, c_vp, C.c_int, c_vp, C.c_int, c_vp, c_vp, C.c_int, C.c_int, c_i64]),
    "rb_ri_dp": (C.c_int, [c_vp, c_vp, c_vp, c_vp, C.c_int, C.c_int]),
    "rb_ri_j": (C.c_int, [c_vp, c_vp, c_vp, c_vp, C.c_int, C.c_int]),
    "rb_ri_dp_j": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, C.c_int, C.c_int]),
    "rb_ri_k": (C.c_int, [c_vp, c_vp, c_vp, C.c_int, c_vp, C.c_int, C.c_int]),
    "rb_ri_iajb": (C.c_int, [c_vp, C.c_int, c_vp, c_i64] + [C.c_int] * 6 + [c_vp, c_i64] + [C.c_int] * 6
                   + [C.c_double, c_vp, c_i64]),
    "rb_host_ri_iajb": (C.c_int, [C.c_int, c_vp] + [C.c_int] * 6 + [c_vp] + [C.c_int] * 6 + [c_vp]),
    "rb_ri_mo_pq": (C.c_int, [c_vp, c_vp, c_i64, C.c_int, c_vp, c_i64, C.c_int] + [C.c_int] * 6
                    + [c_vp, C.c_double, c_vp, c_i64]),
    "rb_host_ri_mo_pq": (C.c_int, [c_vp] + [C.c_int] * 7 + [c_vp, c_vp]),
    "rb_ri_mo_pq_peers": (C.c_int, [c_vp, C.c_int, C.c_int, C.POINTER(c_vp), c_i64, c_ip, c_i64, c_vp, c_vp, c_i64,
                                    C.POINTER(c_i64)]),
    "rb_dsyev": (C.c_int, [c_vp, C.c_char, C.c_char, C.c_int, c_vp, c_i64, c_vp, c_vp, c_i64]),
    "rb_dspev": (C.c_int, [c_vp, C.c_int, c_vp, c_vp, c_vp, c_i64]),
    "rb_dspgv": (C.c_int, [c_vp, C.c_int, c_vp, c_vp, C.c_int, c_vp, c_vp, c_i64]),
    "rb_matrix_power": (C.c_int, [c_vp, C.c_int, c_vp, c_i64, C.c_double, C.c_double, c_vp, c_i64, c_ip]),
    "rb_host_dsyev": (C.c_int, [C.c_char, C.c_int, c_vp, c_vp, c_vp]),
    "rb_host_dspevx": (C.c_int, [C.c_int, c_vp, c_vp, c_vp, c_ip]),
    "rb_host_dspgvx": (C.c_int, [C.c_int, c_vp, c_vp, C.c_int, c_vp, c_vp]),
    "rb_host_power": (C.c_int, [C.c_int, c_vp, C.c_double, C.c_double, c_vp, c_ip]),
    "rb_special_dgemm_01_peers": (C.c_int, [c_vp, C.c_int, C.c_int, C.POINTER(c_vp), c_i64, c_ip, C.POINTER(c_i64), c_vp, c_i64,
                                            C.c_double, C.c_double, c_vp]),
    "rb_special_dgemm_01": (C.c_int, [c_vp, c_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                      c_vp, c_i64, C.c_int, C.c_double, C.c_double]),
    "rb_einsum_ij_j": (C.c_int, [c_vp, c_vp, c_i64, c_vp, c_vp, c_i64, c_i64, c_i64]),
    "rb_einsum_ip_ip": (C.c_int, [c_vp, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_i64]),
    "rb_einsum_i_j": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64]),
    "rb_host_einsum": (C.c_int, [C.c_int, c_vp, c_vp, c_vp, c_i64, c_i64]),
    "rb_pack_upper": (C.c_int, [c_vp, c_vp, c_i64, c_vp]),
    "rb_unpack_upper": (C.c_int, [c_vp, c_vp, c_i64, c_vp]),
    "rb_ri_pack_symm": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_vp]),
    "rb_copy_mm": (C.c_int, [c_vp, C.c_int, C.c_int, c_vp] + [C.c_int] * 4 + [c_vp] + [C.c_int] * 4),
    "rb_copy_mr": (C.c_int, [c_vp, C.c_int, C.c_int, c_vp] + [C.c_int] * 4 + [c_vp] + [C.c_int] * 7),
    "rb_copy_rm": (C.c_int, [c_vp, C.c_int, C.c_int, c_vp] + [C.c_int] * 7 + [c_vp] + [C.c_int] * 4),
    "rb_copy_rr": (C.c_int, [c_vp, C.c_int, C.c_int, C.c_int, c_vp] + [C.c_int] * 6 + [c_vp] + [C.c_int] * 6),
    "rb_ri_transpose": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, C.c_int, c_vp]),
    "rb_matrix_transpose": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_vp]),
    "rb_self_scaled_add": (C.c_int, [c_vp, c_vp, c_vp, C.c_double, c_i64]),
    "rb_self_general_add": (C.c_int, [c_vp, c_vp, c_vp, C.c_double, C.c_double, c_i64]),
    "rb_self_multiple": (C.c_int, [c_vp, c_vp, C.c_double, c_i64]),
    "rb_self_add": (C.c_int, [c_vp, c_vp, c_vp, c_i64]),
    "rb_self_sub": (C.c_int, [c_vp, c_vp, c_vp, c_i64]),
    "rb_fill_linear": (C.c_int, [c_vp, c_vp, c_i64, C.c_uint64, C.c_uint64, C.c_double]),
    "rb_fill_ri3ao_symm": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, C.c_uint64, C.c_double]),
    "rb_gemm_plan_splits": (c_i64, [c_i64, c_i64, c_i64, c_i64, C.c_int, C.c_int]),
    "rb_ri_plan_chunk": (c_i64, [c_i64, c_i64, c_i64, C.c_int]),
    "rb_fp64_peak_probe": (C.c_int, [c_vp, C.c_int, C.c_int, c_dp, c_dp]),
    "rb_hbm_copy_probe": (C.c_int, [c_vp, c_i64, C.c_int, c_dp]),
    "rb_pcie_probe": (C.c_int, [c_vp, C.c_int, c_i64, c_i64, C.c_int, c_dp]),
    "rb_bind_host_to_device_numa": (C.c_int, [C.c_int, c_ip]),
    "rb_host_trim": (C.c_int, []),
    # collectives (NCCL bound at run time)
    "rb_comm_nccl_version": (C.c_int, [c_ip, C.c_char_p, C.c_int]),
    "rb_comm_unique_id": (C.c_int, [c_vp]),
    "rb_comm_init_rank": (C.c_int, [c_vp, C.c_int, C.c_int, c_vp]),
    "rb_comm_init_all": (C.c_int, [C.POINTER(c_vp), C.c_int]),
    "rb_comm_destroy": (C.c_int, [c_vp]),
    "rb_comm_rank": (C.c_int, [c_vp]),
    "rb_comm_world": (C.c_int, [c_vp]),
    "rb_comm_group_start": (C.c_int, []),
    "rb_comm_group_end": (C.c_int, []),
    "rb_allreduce_sum": (C.c_int, [c_vp, c_vp, c_i64]),
    "rb_allgather_shards": (C.c_int, [c_vp, c_vp, c_vp, c_i64]),
    "rb_ri_j_allreduce": (C.c_int, [c_vp, c_vp, c_vp, c_vp, C.c_int, C.c_int]),
    "rb_ri_k_allreduce": (C.c_int, [c_vp, c_vp, c_vp, C.c_int, c_vp, C.c_int, C.c_int]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here == the .so does not export what the header declares
    _fn.restype = _res
    _fn.argtypes = _args


def last_error() -> str:
    msg = lib.rb_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(status: int, what: str = "") -> None:
    if status != 0:
        raise RestB200Error(f"{what or 'librest_b200'} failed (status {status}): {last_error()}")


def ch(c: str) -> bytes:
    """single ASCII character -> c_char argument"""
    return c.encode("ascii")
