//! Rust side of the drop-in.  The reference's `src/external_libs/ffi_restmatr.rs` stays AS IS: librest_b200.so exports the
//! same seven Fortran-ABI symbols (`ri_ao2mo_f_`, `general_dgemm_f_`, `special_dgemm_f_01_`, `copy_mm_`, `copy_mr_`,
//! `copy_rm_`, `copy_rr_`), so `RIFull::ao2mo`, `_dgemm`, `copy_from_*` need no source change -- only the link line.
//! This crate adds
//!   * `ffi`      every symbol of include/rest_b200.h (generated from the header, checked against it by a CPU test),
//!   * `blas_gpu` stand-ins with the call shape of `blas::{dgemm, dsyrk, dgemv, dsymm}`,
//!   * `tensors`  `RIFull` / `MatrixFull` / `MatrixUpper` with the reference's fields and the hot-path methods bound to the GPU,
//!   * `device`   the device-resident, P-sharded API (ri3ao stays in HBM; NCCL all-reduce of J / K inside the library).
//! See INTEGRATION.md.  No Rust toolchain exists in the image this was written in: the declarations are verified against
//! the header and the exported symbols by tests/test_rust_ffi_matches_header.py instead of by rustc.
pub mod ffi;
pub mod blas_gpu;
pub mod tensors;
pub mod device;
