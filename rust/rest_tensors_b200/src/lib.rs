//! Rust side of the drop-in.  The reference's `src/external_libs/ffi_restmatr.rs` stays AS IS: librest_b200.so exports the
//! same seven Fortran-ABI symbols (`ri_ao2mo_f_`, `general_dgemm_f_`, `special_dgemm_f_01_`, `copy_mm_`, `copy_mr_`,
//! `copy_rm_`, `copy_rr_`), so `RIFull::ao2mo`, `_dgemm`, `copy_from_*` need no source change -- only the link line.
//! This crate adds (a) the `rb_host_*` entry points that stand in for the `blas::{dgemm,dsyrk,dgemv,dsymm}` calls and
//! the pure-Rust pack/unpack/transposes, and (b) the device-resident API.  See INTEGRATION.md.
pub mod ffi;
pub mod blas_gpu;
