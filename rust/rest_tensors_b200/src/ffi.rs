//! Raw declarations of include/rest_b200.h (status-returning part).  0 == RB_OK.
use std::ffi::{c_char, c_double, c_int, c_void};

#[repr(C)]
pub struct RbCtx { _private: [u8; 0] }

extern "C" {
    pub fn rb_last_error() -> *const c_char;
    pub fn rb_ctx_create(device: c_int, out: *mut *mut RbCtx) -> c_int;
    pub fn rb_ctx_destroy(ctx: *mut RbCtx) -> c_int;
    pub fn rb_ctx_sync(ctx: *mut RbCtx) -> c_int;
    /// bind the calling thread to the CPUs next to `device` before allocating pinned ri3ao / ri3mo (multi-socket hosts)
    pub fn rb_bind_host_to_device_numa(device: c_int, node_out: *mut c_int) -> c_int;
    pub fn rb_dev_alloc(ctx: *mut RbCtx, bytes: i64, out: *mut *mut c_void) -> c_int;
    pub fn rb_dev_free(ctx: *mut RbCtx, p: *mut c_void) -> c_int;
    /// drop the cached device staging / pinned bounce blocks of the host-pointer entry points
    pub fn rb_host_trim() -> c_int;
    pub fn rb_host_alloc_pinned(bytes: i64, out: *mut *mut c_void) -> c_int;
    pub fn rb_host_free_pinned(p: *mut c_void) -> c_int;
    pub fn rb_memcpy_h2d(ctx: *mut RbCtx, dst: *mut c_void, src: *const c_void, bytes: i64) -> c_int;
    pub fn rb_memcpy_d2h(ctx: *mut RbCtx, dst: *mut c_void, src: *const c_void, bytes: i64) -> c_int;

    // host-pointer BLAS / layout wrappers
    pub fn rb_host_dgemm(ta: c_char, tb: c_char, m: c_int, n: c_int, k: c_int, alpha: c_double, a: *const c_double,
                         lda: c_int, b: *const c_double, ldb: c_int, beta: c_double, c: *mut c_double, ldc: c_int) -> c_int;
    pub fn rb_host_dsyrk(uplo: c_char, trans: c_char, n: c_int, k: c_int, alpha: c_double, a: *const c_double, lda: c_int,
                         beta: c_double, c: *mut c_double, ldc: c_int) -> c_int;
    pub fn rb_host_dgemv(trans: c_char, m: c_int, n: c_int, alpha: c_double, a: *const c_double, lda: c_int,
                         x: *const c_double, incx: c_int, beta: c_double, y: *mut c_double, incy: c_int) -> c_int;
    pub fn rb_host_dsymm(side: c_char, uplo: c_char, m: c_int, n: c_int, alpha: c_double, a: *const c_double, lda: c_int,
                         b: *const c_double, ldb: c_int, beta: c_double, c: *mut c_double, ldc: c_int) -> c_int;
    pub fn rb_host_to_matrixupper(full: *const c_double, n: i64, packed: *mut c_double) -> c_int;
    pub fn rb_host_to_matrixfull(packed: *const c_double, len: i64, full: *mut c_double) -> c_int;
    pub fn rb_host_ri_pack_symm(ri: *const c_double, nao: i64, naux: i64, out: *mut c_double) -> c_int;
    pub fn rb_host_ri_transpose(inp: *const c_double, i: i64, j: i64, k: i64, which: c_int, out: *mut c_double) -> c_int;
    pub fn rb_host_matrix_transpose(inp: *const c_double, rows: i64, cols: i64, out: *mut c_double) -> c_int;
    pub fn rb_host_axpy(op: c_int, c: *mut c_double, p: *const c_double, a: c_double, b: c_double, n: i64) -> c_int;
    pub fn rb_host_ri_ao2mo(cl: *const c_double, nl: c_int, cr: *const c_double, nr: c_int, ri3ao: *const c_double,
                            out: *mut c_double, nb: c_int, nx: c_int) -> c_int;
    pub fn rb_host_ri_ao2mo_jk(cl: *const c_double, nl: c_int, cr: *const c_double, nr: c_int, ri3ao: *const c_double,
                               ri3mo: *mut c_double, nb: c_int, nx: c_int, dm: *const c_double, ct: *const c_double,
                               no: c_int, d: *mut c_double, j: *mut c_double, k: *mut c_double) -> c_int;
    /// which = 1 "ij,j->ij", 2 "ip,ip->p", 3 "i,j->ij" (matrix_blas_lapack.rs:1273-1387)
    pub fn rb_host_einsum(which: c_int, a: *const c_double, b: *const c_double, out: *mut c_double, ni: i64, nj: i64) -> c_int;
    pub fn rb_host_ri_dp(ri3ao: *const c_double, dm: *const c_double, d: *mut c_double, nb: c_int, nx: c_int) -> c_int;
    pub fn rb_host_ri_j(ri3ao: *const c_double, d: *const c_double, j: *mut c_double, nb: c_int, nx: c_int) -> c_int;
    pub fn rb_host_ri_k(ri3ao: *const c_double, ct: *const c_double, no: c_int, k: *mut c_double, nb: c_int, nx: c_int) -> c_int;
    /// (ia|jb)-type blocks of host ri3mo tensors [np, nl, nr]; out = dense [lla*rla, llb*rlb]
    pub fn rb_host_ri_iajb(np: c_int, mo_a: *const c_double, nl_a: c_int, nr_a: c_int, l0a: c_int, lla: c_int, r0a: c_int,
                           rla: c_int, mo_b: *const c_double, nl_b: c_int, nr_b: c_int, l0b: c_int, llb: c_int, r0b: c_int,
                           rlb: c_int, out: *mut c_double) -> c_int;
    /// RPA-type out[P,Q] = sum_{(l,r) in box} w[l,r] mo[P,l,r] mo[Q,l,r]; w may be null
    pub fn rb_host_ri_mo_pq(mo: *const c_double, np: c_int, nl: c_int, nr: c_int, l0: c_int, ll: c_int, r0: c_int, rl: c_int,
                            w: *const c_double, out: *mut c_double) -> c_int;
    /// eigen-solvers with the reference's conventions (matrix_blas_lapack.rs:319-352, 599-652, 1075-1147, 2123-2185)
    pub fn rb_host_dsyev(jobz: c_char, n: c_int, a: *const c_double, w: *mut c_double, z: *mut c_double) -> c_int;
    pub fn rb_host_dspevx(n: c_int, ap: *const c_double, w: *mut c_double, z: *mut c_double, n_found: *mut c_int) -> c_int;
    pub fn rb_host_dspgvx(n: c_int, ap: *const c_double, bp: *const c_double, num_orb: c_int, w: *mut c_double, z: *mut c_double) -> c_int;
    pub fn rb_host_power(n: c_int, a: *const c_double, p: c_double, threshold: c_double, out: *mut c_double, n_nonsingular: *mut c_int) -> c_int;

    // device-resident API (device pointers)
    pub fn rb_ri_ao2mo(ctx: *mut RbCtx, cl: *const c_double, nl: c_int, cr: *const c_double, nr: c_int,
                       ri3ao: *const c_double, out: *mut c_double, nb: c_int, nx: c_int, out_ldp: i64) -> c_int;
    pub fn rb_ri_dp(ctx: *mut RbCtx, ri3ao: *const c_double, dm: *const c_double, d: *mut c_double, nb: c_int, nx: c_int) -> c_int;
    pub fn rb_ri_j(ctx: *mut RbCtx, ri3ao: *const c_double, d: *const c_double, j: *mut c_double, nb: c_int, nx: c_int) -> c_int;
    pub fn rb_ri_k(ctx: *mut RbCtx, ri3ao: *const c_double, ct: *const c_double, no: c_int, k: *mut c_double, nb: c_int, nx: c_int) -> c_int;
    pub fn rb_ri_iajb(ctx: *mut RbCtx, np: c_int, mo_a: *const c_double, ldp_a: i64, nl_a: c_int, nr_a: c_int, l0a: c_int,
                      lla: c_int, r0a: c_int, rla: c_int, mo_b: *const c_double, ldp_b: i64, nl_b: c_int, nr_b: c_int,
                      l0b: c_int, llb: c_int, r0b: c_int, rlb: c_int, beta: c_double, out: *mut c_double, ldo: i64) -> c_int;
    pub fn rb_peer_enable(ctx: *mut RbCtx, peer_device: c_int) -> c_int;
    pub fn rb_ipc_export(ctx: *mut RbCtx, dev_ptr: *mut std::ffi::c_void, handle: *mut u8, offset_out: *mut i64) -> c_int;
    pub fn rb_special_dgemm_01_peers(ctx: *mut RbCtx, rank: c_int, world: c_int, shards: *const *const c_double, xy: i64,
                                     np: *const c_int, p_off: *const i64, b: *const c_double, ldb: i64, alpha: c_double,
                                     beta: c_double, out: *mut c_double) -> c_int;
    pub fn rb_ipc_open(ctx: *mut RbCtx, handle: *const u8, out: *mut *mut std::ffi::c_void) -> c_int;
    pub fn rb_ipc_close(ctx: *mut RbCtx, ptr: *mut std::ffi::c_void) -> c_int;
    pub fn rb_ri_mo_pq_peers(ctx: *mut RbCtx, rank: c_int, world: c_int, panels: *const *const c_double, ld: i64,
                             np: *const c_int, cols: i64, w: *const c_double, out: *mut c_double, ldo: i64,
                             q_off: *const i64) -> c_int;
    pub fn rb_ri_mo_pq(ctx: *mut RbCtx, mo_a: *const c_double, ldp_a: i64, np_a: c_int, mo_b: *const c_double, ldp_b: i64,
                       np_b: c_int, nl: c_int, nr: c_int, l0: c_int, ll: c_int, r0: c_int, rl: c_int, w: *const c_double,
                       beta: c_double, out: *mut c_double, ldo: i64) -> c_int;
}

/// Turn a non-zero status into the panic the reference's wrappers raise on bad shapes.
pub fn check(status: c_int, what: &str) {
    if status != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(rb_last_error()) }.to_string_lossy().into_owned();
        panic!("{} failed (status {}): {}", what, status, msg);
    }
}
