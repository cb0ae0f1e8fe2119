//! Stand-ins with the call shape of `blas::{dgemm, dsyrk, dgemv, dsymm}` (crate blas 0.22) as the reference uses them in
//! src/matrix/matrix_blas_lapack.rs:67,229-241,372-375,410,723,757-769.  In that file replace
//!     use blas::{dgemm, dtrmm, dsymm, dsyrk, dgemv};
//! by
//!     use rest_tensors_b200::blas_gpu::{dgemm, dsymm, dsyrk, dgemv};
//! and every wrapper (`_dgemm_full`, `_dsyrk`, `_dgemv`, `_dsymm`, `ddot`, `lapack_dgemm`) runs on the GPU unchanged.
use crate::ffi::*;
use std::ffi::c_char;

#[allow(clippy::too_many_arguments)]
pub unsafe fn dgemm(transa: u8, transb: u8, m: i32, n: i32, k: i32, alpha: f64, a: &[f64], lda: i32, b: &[f64], ldb: i32,
                    beta: f64, c: &mut [f64], ldc: i32) {
    check(rb_host_dgemm(transa as c_char, transb as c_char, m, n, k, alpha, a.as_ptr(), lda, b.as_ptr(), ldb, beta,
                        c.as_mut_ptr(), ldc), "dgemm");
}

#[allow(clippy::too_many_arguments)]
pub unsafe fn dsyrk(uplo: u8, trans: u8, n: i32, k: i32, alpha: f64, a: &[f64], lda: i32, beta: f64, c: &mut [f64], ldc: i32) {
    check(rb_host_dsyrk(uplo as c_char, trans as c_char, n, k, alpha, a.as_ptr(), lda, beta, c.as_mut_ptr(), ldc), "dsyrk");
}

#[allow(clippy::too_many_arguments)]
pub unsafe fn dgemv(trans: u8, m: i32, n: i32, alpha: f64, a: &[f64], lda: i32, x: &[f64], incx: i32, beta: f64,
                    y: &mut [f64], incy: i32) {
    check(rb_host_dgemv(trans as c_char, m, n, alpha, a.as_ptr(), lda, x.as_ptr(), incx, beta, y.as_mut_ptr(), incy), "dgemv");
}

#[allow(clippy::too_many_arguments)]
pub unsafe fn dsymm(side: u8, uplo: u8, m: i32, n: i32, alpha: f64, a: &[f64], lda: i32, b: &[f64], ldb: i32, beta: f64,
                    c: &mut [f64], ldc: i32) {
    check(rb_host_dsymm(side as c_char, uplo as c_char, m, n, alpha, a.as_ptr(), lda, b.as_ptr(), ldb, beta,
                        c.as_mut_ptr(), ldc), "dsymm");
}
