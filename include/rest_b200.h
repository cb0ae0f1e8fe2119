/*
 * rest_b200.h -- C ABI of librest_b200.so, the B200 (sm_100a) implementation of the rest_tensors RI hot path.
 *
 * Two families of entry points:
 *
 *  (1) COMPAT symbols -- exactly the symbols the reference's Rust FFI binds to its Fortran librestmatr
 *      (reference src/external_libs/ffi_restmatr.rs:4-62; linked by name through build.rs:34
 *      `cargo:rustc-link-lib=restmatr`).  gfortran ABI: every argument by pointer, 32-bit ints, no hidden
 *      string-length argument (the Rust side passes none, src/external_libs/mod.rs:70,75).  All data
 *      pointers are HOST pointers; the call uploads, runs the CUDA kernels and downloads before returning.
 *      They return void like the Fortran; on a CUDA failure they print rb_last_error() and abort()
 *      (the same contract as BLAS xerbla).
 *
 *  (2) rb_* API -- status-returning.  `rb_host_*` take HOST pointers and mirror the reference's safe Rust
 *      wrappers around OpenBLAS (src/matrix/matrix_blas_lapack.rs) and its pure-Rust pack/unpack/transposes.
 *      The remaining rb_* take DEVICE pointers (column-major FP64, 8-byte aligned; 16-byte alignment and even
 *      leading dimensions enable the TMA fast path) and run asynchronously on the context's stream, so that
 *      ri3ao can stay resident (and P-sharded) in HBM across SCF iterations.
 *
 * All matrices/tensors are column-major, exactly as in the reference (RIFull: x + y*s0 + z*s0*s1,
 * src/ri.rs:18-40; MatrixFull: i + j*rows, src/matrix/mod.rs:472-480; MatrixUpper: j(j+1)/2+i,
 * src/index.rs:209-226).
 *
 * There is no CPU fallback anywhere in this library: with no usable CUDA device every entry point fails
 * (rb_* return RB_ERR_CUDA, compat symbols abort).
 */
#ifndef REST_B200_H
#define REST_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RB_OK 0
#define RB_ERR_INVALID 1     /* bad argument / shape mismatch (the Rust wrappers panic on these) */
#define RB_ERR_CUDA 2        /* CUDA runtime / driver failure, or no device */
#define RB_ERR_UNSUPPORTED 3 /* valid request this build does not implement */
#define RB_ERR_NOMEM 4

typedef struct rb_ctx rb_ctx;

/* ---- library / context ------------------------------------------------------------------------------- */
int rb_version(void);
const char *rb_last_error(void); /* thread-local message of the last failing call on this thread */
int rb_device_count(void);

int rb_ctx_create(int device, rb_ctx **out);
int rb_ctx_destroy(rb_ctx *ctx);
/* Run subsequent calls on `cuda_stream` (a cudaStream_t, e.g. torch's current stream; NULL = the legacy default
 * stream).  A fresh context runs on its own non-blocking stream; rb_ctx_use_own_stream() goes back to it. */
int rb_ctx_set_stream(rb_ctx *ctx, void *cuda_stream);
int rb_ctx_use_own_stream(rb_ctx *ctx);
int rb_ctx_sync(rb_ctx *ctx);
int rb_ctx_num_sms(rb_ctx *ctx);
/* Number of kernels this context has launched since creation (bench.py's gpu_launches). */
int64_t rb_ctx_launch_count(rb_ctx *ctx);
/* ... of which launches of the bulk-tensor (TMA load -> shared memory -> TMA store) copy / transpose kernels;
 * the tests use it to prove that aligned operands take that path and not the plain-load fallback. */
int64_t rb_ctx_tma_layout_count(rb_ctx *ctx);
/* Select the GEMM implementation: 0 = auto (TMA+DMMA when alignment allows, else generic DMMA), 1 = force generic. */
int rb_ctx_set_gemm_path(rb_ctx *ctx, int path);
/* ---- CUDA-graph recording of a repeated call sequence ------------------------------------------------------------------------
 * The reference's SCF loop (d_P, J, K once per iteration on the same tensors) repeats one call
 * sequence on the same buffers.  Between rb_graph_begin and rb_graph_end the calls issued on `ctx` are recorded instead of run;
 * rb_graph_launch replays the recording with ONE launch (pointers, shapes, scalars and the split-K / stream-K plans are baked in:
 * refresh the CONTENTS of the input buffers between replays, e.g. with rb_memcpy_h2d, which can itself be part of the recording
 * when the host buffer is pinned).  Rules: the context must run on a non-default stream (rb_ctx_use_own_stream / rb_ctx_set_stream);
 * run the sequence once before recording (workspaces get their size on first use and cannot grow while recording:
 * RB_ERR_UNSUPPORTED); calls that need the host inside (eigen-solvers, the peer pipelines, probes, rb_ctx_sync) are refused with
 * RB_ERR_UNSUPPORTED while a recording is open.  The NCCL collectives (rb_allreduce_sum, rb_ri_j/k_allreduce, rb_allgather_shards)
 * CAN be recorded: every rank must record, and later replay, the same sequence.  A recording stays valid until rb_graph_free, also across later,
 * larger calls on the same context. */
int rb_graph_begin(rb_ctx *ctx);
int rb_graph_end(rb_ctx *ctx, void **graph_out);
int rb_graph_launch(rb_ctx *ctx, void *graph);
int64_t rb_graph_kernel_count(void *graph);
int rb_graph_free(rb_ctx *ctx, void *graph);
/* Test hook: fill every internal workspace with the all-ones bit pattern (NaN as a double) on the context's stream, so that a kernel
 * reading workspace memory nobody wrote in the same call yields NaN instead of silently reusing the previous call's data. */
int rb_ctx_poison_workspaces(rb_ctx *ctx);
/* Select the copy / transpose implementation: 0 = 32-byte (LDG/STG.256) and plain-load kernels (default: measured faster),
 * 1 = bulk-tensor kernels (TMA load -> shared memory -> TMA store) for operands TMA can describe. */
int rb_ctx_set_layout_path(rb_ctx *ctx, int path);

/* device memory helpers for hosts without their own allocator (Rust/C++ side) */
int rb_dev_alloc(rb_ctx *ctx, int64_t bytes, void **out);
int rb_dev_free(rb_ctx *ctx, void *p);
int rb_host_alloc_pinned(int64_t bytes, void **out);
int rb_host_free_pinned(void *p);
/* Page-lock / release a buffer the caller owns (a long-lived Vec<f64> such as ri3ao): the host-pointer entry points then stream it at
 * the pinned rate (config C: ~120 ms per pass instead of ~250 ms through the pageable bounce).  Unregister before freeing the memory. */
int rb_host_register(void *p, int64_t bytes);
int rb_host_unregister(void *p);
int rb_memcpy_h2d(rb_ctx *ctx, void *dst, const void *src, int64_t bytes); /* async on the ctx stream */
int rb_memcpy_d2h(rb_ctx *ctx, void *dst, const void *src, int64_t bytes);

/* Peer memory over NVLink / NVSwitch.  Every rb_* device entry point accepts operands that live in another GPU's HBM
 * once the mapping exists (the TMA tensor maps are encoded over plain global addresses):
 *   same process, several devices : rb_peer_enable(ctx, peer_device) once per reading context;
 *   one process per GPU           : the owner exports a device buffer as a 64-byte handle of the allocation that holds
 *                                   it plus the buffer's byte offset inside that allocation (offset_out may be NULL for
 *                                   a block that came straight from rb_dev_alloc), every reader opens the handle
 *                                   (rb_ipc_open enables peer access lazily; add the offset) and closes it when done.
 * Used by rb_ri_mo_pq_peers and rb_special_dgemm_01_peers, which consume other ranks' blocks. */
int rb_peer_enable(rb_ctx *ctx, int peer_device);
int rb_ipc_export(rb_ctx *ctx, void *dev_ptr, unsigned char handle[64], int64_t *offset_out);
int rb_ipc_open(rb_ctx *ctx, const unsigned char handle[64], void **out);
int rb_ipc_close(rb_ctx *ctx, void *ptr);

/* ---- (1) COMPAT: reference src/external_libs/ffi_restmatr.rs:4-62 == restmatr.f90 ---------------------- */

/* ffi_restmatr.rs:5-11 / restmatr.f90:158-194:
 * ri3mo[P + a*nx + b*nx*ns] = sum_mu C[mu,a] sum_nu ri3fn[mu,nu,P] C[nu,b];  ri3mo fully overwritten.
 * The Fortran is only defined for num_states == num_basis (it passes ldc=num_basis for an ns x ns section);
 * this implementation gives the same result there and the natural rectangular result otherwise. */
void ri_ao2mo_f_(const double *eigenvector, const double *ri3fn, double *ri3mo, const int *num_states,
                 const int *num_basis, const int *num_auxbas);

/* ffi_restmatr.rs:13-24 / restmatr.f90:63-107: C[rc,cc] = alpha*op(A[ra,ca])*op(B[rb,cb]) + beta*C[rc,cc];
 * starts are 0-based; elements of C outside the block are untouched; op chars 'N' | 'T'. */
void general_dgemm_f_(const double *matr_a, const int *rows_a, const int *columns_a, const int *start_row_a,
                      const int *len_row_a, const int *start_column_a, const int *len_column_a, const char *opa,
                      const double *matr_b, const int *rows_b, const int *columns_b, const int *start_row_b,
                      const int *len_row_b, const int *start_column_b, const int *len_column_b, const char *opb,
                      double *matr_c, const int *rows_c, const int *columns_c, const int *start_row_c,
                      const int *len_row_c, const int *start_column_c, const int *len_column_c, const double *alpha,
                      const double *beta);

/* ffi_restmatr.rs:25-33 / restmatr.f90:111-154: for every y, in place:
 * T[xr, y, zr] <- alpha * T[xr, y, zr] * B[rb, cb] + beta * T[xr, y, zr]   (needs len_column_b == len_z_a). */
void special_dgemm_f_01_(double *ten3_a, const int *x_a, const int *y_a, const int *z_a, const int *start_x_a,
                         const int *len_x_a, const int *i_y, const int *start_z_a, const int *len_z_a,
                         const double *matr_b, const int *rows_b, const int *columns_b, const int *start_row_b,
                         const int *len_row_b, const int *start_column_b, const int *len_column_b,
                         const double *alpha, const double *beta);

/* ffi_restmatr.rs:35-39 / restmatr.f90:197-212 */
void copy_mm_(const int *x_len, const int *y_len, const double *f_matr, const int *f_x_len, const int *f_y_len,
              const int *f_x_start, const int *f_y_start, double *t_matr, const int *t_x_len, const int *t_y_len,
              const int *t_x_start, const int *t_y_start);
/* ffi_restmatr.rs:41-46 / restmatr.f90:215-238; mod 0: t(x1,x2,x3) 1: t(x1,x3,x2) 2: t(x3,x1,x2); else no-op */
void copy_mr_(const int *x_len, const int *y_len, const double *f_matr, const int *f_x_len, const int *f_y_len,
              const int *f_x_start, const int *f_y_start, double *t_ri, const int *t_x_len, const int *t_y_len,
              const int *t_z_len, const int *t_x_start, const int *t_y_start, const int *t_x3, const int *t_mod);
/* ffi_restmatr.rs:48-53 / restmatr.f90:241-264 */
void copy_rm_(const int *x_len, const int *y_len, const double *f_ri, const int *f_x_len, const int *f_y_len,
              const int *f_z_len, const int *f_x_start, const int *f_y_start, const int *f_x3, const int *f_mod,
              double *t_matr, const int *t_x_len, const int *t_y_len, const int *t_x_start, const int *t_y_start);
/* ffi_restmatr.rs:55-61 / restmatr.f90:266-285 */
void copy_rr_(const int *x_len, const int *y_len, const int *z_len, const double *f_ri, const int *f_x_len,
              const int *f_y_len, const int *f_z_len, const int *f_x_start, const int *f_y_start,
              const int *f_z_start, double *t_ri, const int *t_x_len, const int *t_y_len, const int *t_z_len,
              const int *t_x_start, const int *t_y_start, const int *t_z_start);

/* ---- (2a) HOST-pointer wrappers of the reference's BLAS / layout calls --------------------------------- */
/* These use a process-wide default context on device $REST_B200_DEVICE (default 0) and serialise on a mutex. */

/* blas::dgemm as called from _dgemm_full / lapack_dgemm / ddot (matrix_blas_lapack.rs:180-252,714-774) */
int rb_host_dgemm(char transa, char transb, int m, int n, int k, double alpha, const double *a, int lda,
                  const double *b, int ldb, double beta, double *c, int ldc);
/* blas::dsyrk as called from _dsyrk (matrix_blas_lapack.rs:392-413): only the uplo triangle of C is touched */
int rb_host_dsyrk(char uplo, char trans, int n, int k, double alpha, const double *a, int lda, double beta,
                  double *c, int ldc);
/* blas::dgemv as called from _dgemv (matrix_blas_lapack.rs:38-70) */
int rb_host_dgemv(char trans, int m, int n, double alpha, const double *a, int lda, const double *x, int incx,
                  double beta, double *y, int incy);
/* blas::dsymm as called from _dsymm (matrix_blas_lapack.rs:354-378) */
int rb_host_dsymm(char side, char uplo, int m, int n, double alpha, const double *a, int lda, const double *b,
                  int ldb, double beta, double *c, int ldc);
/* MatrixFull::to_matrixupper (matrixfull.rs:638-646): packed[j(j+1)/2+i] = full[i+j*n], i<=j */
int rb_host_to_matrixupper(const double *full, int64_t n, double *packed);
/* MatrixUpper::to_matrixfull (matrixupper.rs:330-373): len must be triangular (else RB_ERR_INVALID == None) */
int rb_host_to_matrixfull(const double *packed, int64_t len, double *full);
/* RIFull::rifull_to_matfull_symm (ri.rs:297-341) */
int rb_host_ri_pack_symm(const double *ri, int64_t nao, int64_t naux, double *out);
/* RIFull::transpose_{jik,jki,kji,ikj} (ri.rs:227-294); which = 0,1,2,3 */
int rb_host_ri_transpose(const double *in, int64_t i, int64_t j, int64_t k, int which, double *out);
/* MatrixFull::transpose (matrixfull.rs:579-614) */
int rb_host_matrix_transpose(const double *in, int64_t rows, int64_t cols, double *out);
/* Rectangular ao2mo (north-star occ-vir form): out[P + a*nx + b*nx*nl] = sum C_L[mu,a] A[mu,nu,P] C_R[nu,b] */
int rb_host_ri_ao2mo(const double *c_left, int nl, const double *c_right, int nr, const double *ri3ao, double *out,
                     int nb, int nx);
/* Fused streaming step over a HOST ri3ao: each P-chunk is uploaded once and feeds ao2mo (ri3mo != NULL), d_P/J
 * (dm != NULL; d, j may be NULL) and K (ct, k != NULL) -- H2D | DMMA | D2H overlapped on three streams. */
int rb_host_ri_ao2mo_jk(const double *c_left, int nl, const double *c_right, int nr, const double *ri3ao,
                        double *ri3mo, int nb, int nx, const double *dm, const double *ct, int no, double *d, double *j,
                        double *k);
/* Same pass with ONE coefficient matrix c [nb, nmo] on both sides, shipping only the a <= b part of the transformed tensor:
 *   ri3mo_upper[P + nx * (b (b + 1) / 2 + a)] = ri3mo[P, a, b],  0 <= a <= b < nmo   (MatrixUpper's pair index, P fastest).
 * The slabs of ri3ao are symmetric (RIFull built by the reference's 3-centre integral code), so ri3mo[P, a, b] == ri3mo[P, b, a] and
 * nothing is lost, while the device -> host traffic of the pass halves (nmo (nmo + 1) / 2 instead of nmo^2 columns).  For slabs that
 * are not symmetric the result is still exactly the a <= b entries of rb_host_ri_ao2mo_jk's output. */
int rb_host_ri_ao2mo_jk_upper(const double *c, int nmo, const double *ri3ao, double *ri3mo_upper, int nb, int nx,
                              const double *dm, const double *ct, int no, double *d, double *j, double *k);
/* As rb_host_ri_ao2mo_jk_upper, for callers that GUARANTEE symmetric slabs (ri3ao[mu, nu, P] == ri3ao[nu, mu, P], what the reference's
 * 3-centre integrals are): only the mu <= nu part of every slab is uploaded (pitched 3-D copies of 32-column trapezoids, 52-55 % of the
 * bytes) and mirrored in HBM, so the pass moves about half the bytes in BOTH directions.  The strict lower triangle of the host slabs is
 * never read.  ri3mo_upper may be NULL (d_P / J / K only). */
int rb_host_ri_ao2mo_jk_symm(const double *c, int nmo, const double *ri3ao, double *ri3mo_upper, int nb, int nx,
                             const double *dm, const double *ct, int no, double *d, double *j, double *k);
/* axpy family on host buffers (matrix/mod.rs:545-648, ri.rs:345-354, matrixupper.rs:395-420):
 * op 0: c += p*b   1: c = c*a + p*b   2: c *= a   3: c += p   4: c -= p   (unfused mul-then-add, bit-exact) */
int rb_host_axpy(int op, double *c, const double *p, double a, double b, int64_t n);
/* einsum helpers on host buffers: which = 1 "ij,j->ij" (a [ni,nj], b [nj], out [ni,nj]); 2 "ip,ip->p" (a, b [ni,nj],
 * out [nj]); 3 "i,j->ij" (a [ni], b [nj], out [ni,nj]).  Dense column-major operands (lda = ldb = ni). */
int rb_host_einsum(int which, const double *a, const double *b, double *out, int64_t ni, int64_t nj);
/* (ia|jb)-type blocks from dense host tensors moA[np, nl_a, nr_a], moB[np, nl_b, nr_b] (moB may alias moA); out is the
 * dense [lla*rla, llb*rlb] block, overwritten.  See rb_ri_iajb. */
int rb_host_ri_iajb(int np, const double *mo_a, int nl_a, int nr_a, int l0a, int lla, int r0a, int rla,
                    const double *mo_b, int nl_b, int nr_b, int l0b, int llb, int r0b, int rlb, double *out);
/* RPA-type consumer from a dense host ri3mo[np, nl, nr]; out is the dense symmetric [np, np] matrix.  See rb_ri_mo_pq. */
int rb_host_ri_mo_pq(const double *mo, int np, int nl, int nr, int l0, int ll, int r0, int rl, const double *w,
                     double *out);
/* eigen-solvers on host buffers, with the reference's argument conventions: _dsyev / lapack_dsyev (lower triangle read),
 * lapack_dspevx (*n_found = n), lapack_dspgvx / _dspgvx (z [n, num_orb]), _power / lapack_power */
int rb_host_dsyev(char jobz, int n, const double *a, double *w, double *z);
int rb_host_dspevx(int n, const double *ap, double *w, double *z, int *n_found);
int rb_host_dspgvx(int n, const double *ap, const double *bp, int num_orb, double *w, double *z);
int rb_host_power(int n, const double *a, double p, double threshold, double *out, int *n_nonsingular);
/* d_P, J, K with host buffers (composed by REST from _dgemv/_dgemm/_dsyrk) */
int rb_host_ri_dp(const double *ri3ao, const double *dm, double *d, int nb, int nx);
int rb_host_ri_j(const double *ri3ao, const double *d, double *j, int nb, int nx);
int rb_host_ri_k(const double *ri3ao, const double *ct, int no, double *k, int nb, int nx);

/* ---- (2b) DEVICE-pointer API (async on the ctx stream) -------------------------------------------------- */

/* BLAS-3/2 on device buffers; same argument meaning as the Fortran BLAS. */
int rb_dgemm(rb_ctx *ctx, char transa, char transb, int m, int n, int k, double alpha, const double *a, int64_t lda,
             const double *b, int64_t ldb, double beta, double *c, int64_t ldc);
int rb_dgemm_strided_batched(rb_ctx *ctx, char transa, char transb, int m, int n, int k, double alpha,
                             const double *a, int64_t lda, int64_t stride_a, const double *b, int64_t ldb,
                             int64_t stride_b, double beta, double *c, int64_t ldc, int64_t stride_c, int batch);
int rb_dsyrk(rb_ctx *ctx, char uplo, char trans, int n, int k, double alpha, const double *a, int64_t lda,
             double beta, double *c, int64_t ldc);
int rb_dgemv(rb_ctx *ctx, char trans, int m, int n, double alpha, const double *a, int64_t lda, const double *x,
             int incx, double beta, double *y, int incy);
int rb_dsymm(rb_ctx *ctx, char side, char uplo, int m, int n, double alpha, const double *a, int64_t lda,
             const double *b, int64_t ldb, double beta, double *c, int64_t ldc);

/* RI contractions over the local slabs ri3ao[nb, nb, nx] (one rank's P-shard).
 * ao2mo: out[P + a*out_ldp + b*out_ldp*nl], P in [0,nx): out_ldp >= nx lets a rank write its rows of a larger
 *        (e.g. global-naux) tensor.  c_left [nb,nl], c_right [nb,nr]; square reference semantics: both = C, nl=nr=ns. */
int rb_ri_ao2mo(rb_ctx *ctx, const double *c_left, int nl, const double *c_right, int nr, const double *ri3ao,
                double *out, int nb, int nx, int64_t out_ldp);
/* d[P] = sum_{mu,nu} ri3ao[mu,nu,P]*dm[mu,nu] */
int rb_ri_dp(rb_ctx *ctx, const double *ri3ao, const double *dm, double *d, int nb, int nx);
/* j[mu,nu] = sum_P ri3ao[mu,nu,P]*d[P]  (this rank's partial sum; all-reduce across ranks) */
int rb_ri_j(rb_ctx *ctx, const double *ri3ao, const double *d, double *j, int nb, int nx);
/* d_P and J from ONE read of ri3ao (the reference makes two passes, ri.rs + _dgemv 'T' / 'N'): one persistent cooperative kernel, every
 * CTA owns a fixed ij range in all slabs (its D and J entries stay in registers), the slabs stream through a shared-memory ring of 1-D
 * bulk copies, the per-slab dot products are combined across CTAs in a fixed order (deterministic, no FP64 atomics) and J is updated
 * from the copy still in shared memory.  Same results as rb_ri_dp followed by rb_ri_j up to the summation order inside d_P (1e-15).
 * The single-pass kernel runs where it was measured to win (nb ~ 550 .. 950, tensors >= 256 MB: 1.01 vs 1.46 ms at config C, 0.57 vs
 * 0.84 ms at nb = 900); elsewhere (short runs, nb > ~1100, odd nb) this call is rb_ri_dp followed by rb_ri_j.  REST_B200_DPJ_FUSED=0 / 1
 * forces the two passes / the single pass wherever it can run. */
int rb_ri_dp_j(rb_ctx *ctx, const double *ri3ao, const double *dm, double *d, double *j, int nb, int nx);
/* k = sum_P (A_P ct)(A_P ct)^T, ct [nb,no]; full symmetric nb x nb (both triangles written); partial per rank */
int rb_ri_k(rb_ctx *ctx, const double *ri3ao, const double *ct, int no, double *k, int nb, int nx);
/* in-place slab x matrix of restmatr.f90:111-154 on device buffers */
int rb_special_dgemm_01(rb_ctx *ctx, double *ten3, int x_a, int y_a, int z_a, int start_x, int len_x, int start_z,
                        int len_z, const double *b, int64_t ldb, int len_col_b, double alpha, double beta);

/* (ia|jb)-type consumers of ri3mo (the P-fastest layout of src/ri.rs:381-386 exists for them):
 *   out[(l-l0a) + (r-r0a)*lla + ((l'-l0b) + (r'-r0b)*llb)*ldo] = beta*out + sum_{P<np} moA[P,l,r] * moB[P,l',r']
 * with mo[P + l*ldp + r*ldp*nl].  moA == moB with identical boxes and beta == 0 runs as SYRK + mirror (both triangles
 * written, bitwise symmetric); with beta != 0 the full product is formed, so out need not be symmetric.  One rank's
 * partial sum over its local P rows; all-reduce `out` across P-sharded ranks. */
int rb_ri_iajb(rb_ctx *ctx, int np, const double *mo_a, int64_t ldp_a, int nl_a, int nr_a, int l0a, int lla, int r0a,
               int rla, const double *mo_b, int64_t ldp_b, int nl_b, int nr_b, int l0b, int llb, int r0b, int rlb,
               double beta, double *out, int64_t ldo);

/* RPA-type consumer of ri3mo: contraction over the MO pairs of a box, output in the auxiliary basis,
 *   out[P + Q*ldo] = beta*out + sum_{(l,r) in box} w[(l-l0) + (r-r0)*ll] * moA[P,l,r] * moB[Q,l,r]      (w NULL: ones)
 * moA (np_a rows, pitch ldp_a) and moB (np_b rows, pitch ldp_b) are row blocks (P-shards) of [.., nl, nr] tensors;
 * moA == moB with beta == 0 is computed as the upper triangle + mirror.  'N','T' DMMA GEMM with K = ll*rl. */
int rb_ri_mo_pq(rb_ctx *ctx, const double *mo_a, int64_t ldp_a, int np_a, const double *mo_b, int64_t ldp_b, int np_b,
                int nl, int nr, int l0, int ll, int r0, int rl, const double *w, double beta, double *out, int64_t ldo);

/* P-sharded form of rb_ri_mo_pq, all-gather -> GEMM as ONE pipeline over peer memory: panels[s] = rank s's dense box
 * panel [np[s] rows, pitch ld, cols columns] as seen from this rank (rb_ipc_open mapping, or a local pointer for
 * s == rank); out[:, q_off[s] .. q_off[s] + np[s]) = sum_c w[c] * own[:, c] * panels[s][:, c] for every s.  The peers'
 * panels are pulled by the copy engines on a second stream, one peer ahead of the DMMA GEMM that consumes them, so the
 * NVLink transfer overlaps the math.  Callers synchronise the ranks before (panels complete) and after (panels free). */
int rb_ri_mo_pq_peers(rb_ctx *ctx, int rank, int world, const double *const *panels, int64_t ld, const int *np,
                      int64_t cols, const double *w, double *out, int64_t ldo, const int64_t *q_off);

/* P-sharded form of special_dgemm_f_01 (restmatr.f90:111-154 with the full x range, every y, and z = the auxiliary index):
 *   out[xy, P'] = alpha * sum_P T[xy, P] * B[P, P'] + beta * T[xy, P']   for P' in this rank's shard, P over ALL ranks,
 * T viewed as the matrix [xy = X*Y, naux] whose column block P_s lives on rank s (shards[s], as seen from this rank:
 * rb_ipc_open mapping or local pointer).  The contraction runs over the sharded index, so every rank needs every
 * other rank's shard: row chunks of the peers' shards are pulled by the copy engines over NVLink on a second stream,
 * one chunk ahead of the DMMA GEMM that accumulates them into `out` (a separate buffer: peers are still reading T).
 * b = the full [naux, naux] matrix (ldb >= naux), np[s] / p_off[s] = columns and first column of rank s.
 * Callers synchronise the ranks before (shards complete) and after (then `out` may replace the shard). */
int rb_special_dgemm_01_peers(rb_ctx *ctx, int rank, int world, const double *const *shards, int64_t xy, const int *np,
                              const int64_t *p_off, const double *b, int64_t ldb, double alpha, double beta, double *out);

/* ---- collectives of the P-sharded path -----------------------------------------------------------------
 * The reference has no communication layer; the partition is its iter_auxbas(P_lo..P_hi) (src/ri.rs:190-198) and the only
 * exchanges are ONE all-reduce(sum, f64) each for J and K plus an all-gather of d_P.  They are NCCL calls (NVLink 5 /
 * NVSwitch) on the context's stream.  NCCL is bound at run time: $REST_B200_NCCL_LIB, else a libnccl.so.2 the process
 * already carries (PyTorch's), else the loader's search path; RB_ERR_UNSUPPORTED if none is found.
 *   one process (or thread) per GPU : rank 0 calls rb_comm_unique_id and hands the 128 bytes to every rank by the host's
 *                                     own means (pipe, file, MPI, a torch store); every rank calls rb_comm_init_rank.
 *   one process, n contexts         : rb_comm_init_all; bracket each step's per-context collective calls with
 *                                     rb_comm_group_start / rb_comm_group_end (NCCL group semantics).
 * A context without a communicator is a world of one: the collectives are no-ops and the *_allreduce forms equal the
 * plain ones.  rb_ctx_destroy destroys the communicator. */
int rb_comm_nccl_version(int *version_out, char *path_out, int path_len);
int rb_comm_unique_id(unsigned char id[128]);
int rb_comm_init_rank(rb_ctx *ctx, int rank, int world, const unsigned char id[128]);
int rb_comm_init_all(rb_ctx *const *ctxs, int n);
int rb_comm_destroy(rb_ctx *ctx);
int rb_comm_rank(rb_ctx *ctx);
int rb_comm_world(rb_ctx *ctx);
int rb_comm_group_start(void);
int rb_comm_group_end(void);
/* in-place sum over the ranks, asynchronous on the context's stream */
int rb_allreduce_sum(rb_ctx *ctx, double *buf, int64_t n);
/* full[0..naux) on every rank from the per-rank pieces local[floor(r*naux/G) .. floor((r+1)*naux/G)) (d_P) */
int rb_allgather_shards(rb_ctx *ctx, const double *local, double *full, int64_t naux);
/* rb_ri_j / rb_ri_k followed by the all-reduce: J and K complete on every rank */
int rb_ri_j_allreduce(rb_ctx *ctx, const double *ri3ao, const double *d, double *j, int nb, int nx);
int rb_ri_k_allreduce(rb_ctx *ctx, const double *ri3ao, const double *ct, int no, double *k, int nb, int nx);

/* einsum helpers (matrix_blas_lapack.rs:1273-1387, matrix/einsum.rs) on device buffers:
 * "ij,j->ij" and "i,j->ij" are one multiply per element (bit-exact), "ip,ip->p" is a column dot (1e-10). */
int rb_einsum_ij_j(rb_ctx *ctx, const double *a, int64_t lda, const double *b, double *out, int64_t ldo, int64_t ni,
                   int64_t nj);
int rb_einsum_ip_ip(rb_ctx *ctx, const double *a, int64_t lda, const double *b, int64_t ldb, double *out, int64_t ni,
                    int64_t np);
int rb_einsum_i_j(rb_ctx *ctx, const double *a, const double *b, double *out, int64_t ni, int64_t nj);

/* ERIFold4 (src/eri.rs:170-373): four-index integrals with both pairs folded, a column-major
 * [size0 = npair_ij, size1 = npair_kl] matrix (leading dimension ld), packed pair index j(j+1)/2 + i, i <= j.  The chunk scatter
 * moves a dense shell-quartet block buf[li, lj, lk, ll] (first index fastest) for i in i0..i0+li, ... into the folded tensor:
 *   mode 0 = chunk_copy_from_local_erifull (eri.rs:266-305): elements with k <= l and i <= j
 *   mode 1 = chunk_copy_from_a_full_vector (eri.rs:308-372): k <= l and (i0 < j0: every (i, j); i0 == j0: local ii <= jj;
 *            i0 > j0: nothing)
 * Bit-exact.  RB_ERR_INVALID when the block reaches outside the tensor (the reference panics on the slice). */
int rb_erifold4_chunk_copy(rb_ctx *ctx, double *eri, int64_t size0, int64_t size1, int64_t ld, int i0, int li, int j0, int lj,
                           int k0, int lk, int l0, int ll, const double *buf, int mode);
/* the same on a HOST tensor (dense, ld = size0): only the window of rows the block touches crosses PCIe */
int rb_host_erifold4_chunk_copy(double *eri, int64_t size0, int64_t size1, int i0, int li, int j0, int lj, int k0, int lk, int l0,
                                int ll, const double *buf, int mode);

/* Symmetric eigen-solvers (the reference's LAPACK wrappers _dsyev / lapack_dspevx / lapack_dspgvx /
 * _power, matrix_blas_lapack.rs:319-352, 599-652, 1004-1147, 2123-2185) as a parallel one-sided Jacobi method on device
 * buffers.  Eigenvalues ascending; eigenvectors in the columns of z, normalised, largest component positive (LAPACK
 * leaves the sign open).  These calls synchronise the context's stream (the sweep count is data dependent).
 *   rb_dsyev : a [n, n] (the `uplo` triangle is read), w [n], z [n, n] (jobz 'V') or NULL ('N')
 *   rb_dspev : packed upper ap [n(n+1)/2]
 *   rb_dspgv : A x = lambda B x with packed upper ap, bp (B positive definite, else RB_ERR_INVALID); the m lowest
 *              pairs: w [m], z [n, m] with z^T B z = I
 *   rb_matrix_power : out = sum over eigenvalues lambda_i >= threshold of lambda_i^p v_i v_i^T (lower triangle of a is
 *              read, like the reference's dsyev('L')); *n_nonsingular = number of eigenvalues kept */
int rb_dsyev(rb_ctx *ctx, char jobz, char uplo, int n, const double *a, int64_t lda, double *w, double *z, int64_t ldz);
int rb_dspev(rb_ctx *ctx, int n, const double *ap, double *w, double *z, int64_t ldz);
int rb_dspgv(rb_ctx *ctx, int n, const double *ap, const double *bp, int m, double *w, double *z, int64_t ldz);
int rb_matrix_power(rb_ctx *ctx, int n, const double *a, int64_t lda, double p, double threshold, double *out, int64_t ldo,
                    int *n_nonsingular);

/* pack / unpack (bit-exact) */
int rb_pack_upper(rb_ctx *ctx, const double *full, int64_t n, double *packed);
int rb_unpack_upper(rb_ctx *ctx, const double *packed, int64_t n, double *full);
int rb_ri_pack_symm(rb_ctx *ctx, const double *ri, int64_t nao, int64_t naux, double *out);

/* strided sub-box copies (bit-exact); same argument order as copy_mm_/copy_mr_/copy_rm_/copy_rr_ but by value */
int rb_copy_mm(rb_ctx *ctx, int x_len, int y_len, const double *f, int f_x_len, int f_y_len, int f_x_start,
               int f_y_start, double *t, int t_x_len, int t_y_len, int t_x_start, int t_y_start);
int rb_copy_mr(rb_ctx *ctx, int x_len, int y_len, const double *f, int f_x_len, int f_y_len, int f_x_start,
               int f_y_start, double *t, int t_x_len, int t_y_len, int t_z_len, int t_x_start, int t_y_start,
               int t_x3, int mod);
int rb_copy_rm(rb_ctx *ctx, int x_len, int y_len, const double *f, int f_x_len, int f_y_len, int f_z_len,
               int f_x_start, int f_y_start, int f_x3, int mod, double *t, int t_x_len, int t_y_len, int t_x_start,
               int t_y_start);
int rb_copy_rr(rb_ctx *ctx, int x_len, int y_len, int z_len, const double *f, int f_x_len, int f_y_len, int f_z_len,
               int f_x_start, int f_y_start, int f_z_start, double *t, int t_x_len, int t_y_len, int t_z_len,
               int t_x_start, int t_y_start, int t_z_start);

/* transposes (bit-exact) */
int rb_ri_transpose(rb_ctx *ctx, const double *in, int64_t i, int64_t j, int64_t k, int which, double *out);
int rb_matrix_transpose(rb_ctx *ctx, const double *in, int64_t rows, int64_t cols, double *out);

/* axpy family (matrix/mod.rs:545-648, ri.rs:345-354); unfused mul-then-add, bit-exact vs the Rust loops */
int rb_self_scaled_add(rb_ctx *ctx, double *c, const double *p, double b, int64_t n);               /* c += p*b   */
int rb_self_general_add(rb_ctx *ctx, double *c, const double *p, double a, double b, int64_t n);    /* c = c*a+p*b */
int rb_self_multiple(rb_ctx *ctx, double *c, double a, int64_t n);                                  /* c *= a     */
int rb_self_add(rb_ctx *ctx, double *c, const double *p, int64_t n);                                /* c += p     */
int rb_self_sub(rb_ctx *ctx, double *c, const double *p, int64_t n);                                /* c -= p     */

/* counter-based synthetic inputs generated in HBM; bit-identical to oracle/rest_oracle.c */
int rb_fill_linear(rb_ctx *ctx, double *v, int64_t n, uint64_t seed, uint64_t idx0, double scale);
int rb_fill_ri3ao_symm(rb_ctx *ctx, double *a, int64_t nb, int64_t p_lo, int64_t p_hi, uint64_t seed, double scale);

/* Host-side planners, exported so that they can be tested without a GPU (pure arithmetic, no device call):
 *   rb_gemm_plan_splits : split-K count the TMA+DMMA GEMM uses for (m, n, k, batch, tri) on a chip of num_sms SMs
 *   rb_ri_plan_chunk    : P-chunk length of the RI contractions for nx slabs of bytes_per_slab workspace each under a
 *                         workspace budget (tile_m != 0: P is the M index of a GEMM tile, cheapest tiling wins) */
int64_t rb_gemm_plan_splits(int64_t m, int64_t n, int64_t k, int64_t batch, int tri, int num_sms);
/* 1 when the same product would run as stream-K (every CTA an equal share of the COST of the tile x k-step space, pieces summed in a
 * fixed order) instead of a uniform split; costs_out (NULL or 4 doubles): estimated makespans of the uniform split and of stream-K
 * in full-tile k steps, the CTA count and the largest piece count of the stream-K plan */
int rb_gemm_plan_stream_k(int64_t m, int64_t n, int64_t k, int64_t batch, int tri, int num_sms, double *costs_out);
/* the stream-K partition itself (for tests): CTA c works on the steps from (cta_tile[c], cta_step[c]) up to (cta_tile[c+1], cta_step[c+1]);
 * tile t is covered by tile_pieces[t] consecutive CTAs starting with tile_first[t].  Arrays of >= 161 / 161 / 592 / 592 entries.
 * Returns the CTA count, 0 when the shape does not qualify. */
int rb_gemm_stream_k_tables(int64_t m, int64_t n, int64_t k, int64_t batch, int tri, int num_sms, unsigned *cta_step,
                            unsigned short *cta_tile, unsigned char *tile_first, unsigned char *tile_pieces);
int64_t rb_ri_plan_chunk(int64_t nx, int64_t bytes_per_slab, int64_t budget_bytes, int tile_m);

/* FP64 pipe micro-benchmarks used by bench.py for the roofline denominator: returns achieved TFLOP/s of a
 * register-resident DMMA (kind=0) or DFMA (kind=1) loop over the whole chip, timed with CUDA events. */
int rb_fp64_peak_probe(rb_ctx *ctx, int kind, int iters, double *tflops_out, double *ms_out);
/* Stream copy probe (read+write bytes / time) for the HBM denominator cross-check. */
int rb_hbm_copy_probe(rb_ctx *ctx, int64_t bytes, int iters, double *gbs_out);

/* PCIe probe for the host-pointer paths: mode 0 H2D, 1 D2H, 2 D2H 2-D (rows of `width` bytes), 3 H2D+D2H duplex
 * (sum of both directions), 4 D2H by kernel stores into mapped pinned memory, 5 like 4 with rows of `width` bytes. */
int rb_pcie_probe(rb_ctx *ctx, int mode, int64_t bytes, int64_t width, int iters, double *gbs_out);

/* The host-pointer entry points keep their device staging blocks (<= 24 GB) and the pinned bounce blocks used for
 * pageable caller buffers (<= 8 GB) cached between calls; this releases them. */
int rb_host_trim(void);

/* Host placement helper for the host-pointer paths on multi-socket boxes: restrict the calling thread (and the
 * threads / allocations it makes afterwards) to the CPUs of the NUMA node `device` hangs off, so that pinned buffers
 * are first-touched next to the GPU's PCIe root and its DMA does not cross the socket interconnect.  Call it once
 * per rank before allocating ri3ao / ri3mo.  Writes the node (or -1 when the platform reports none) to *node_out.
 * Returns RB_OK also when there is nothing to bind to (single-node hosts). */
int rb_bind_host_to_device_numa(int device, int *node_out);

#ifdef __cplusplus
}
#endif
#endif /* REST_B200_H */
