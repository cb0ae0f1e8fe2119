// rb_einsum.cu -- the Vxc einsum helpers of the reference: src/matrix/matrix_blas_lapack.rs:1273-1387
// and src/matrix/einsum.rs.  All HBM-bound:
//   "ij,j->ij"  out[i,j] = a[i,j] * b[j]            one multiply per element  -> bit-exact
//   "ip,ip->p"  out[p]   = sum_i a[i,p] * b[i,p]    column dots (the d_P access pattern with two streamed operands)
//   "i,j->ij"   out[i,j] = a[i] * b[j]              outer product             -> bit-exact
// ("ij,jk->ik" is _dgemm_full.)
#include "rb_common.cuh"

__global__ void __launch_bounds__(256) rb_einsum_ij_j_kernel(const double *__restrict__ a, i64 lda,
                                                             const double *__restrict__ b, double *__restrict__ out,
                                                             i64 ldo, i64 ni, i64 nj)
{
    // one column per blockIdx.y step, threads along i with 4 loads in flight
    for (i64 j = blockIdx.y; j < nj; j += gridDim.y) {
        const double bj = b[j];
        const double *ac = a + j * lda;
        double *oc = out + j * ldo;
        const i64 stride = (i64)gridDim.x * blockDim.x;
        i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
        for (; i + 3 * stride < ni; i += 4 * stride) {
            const double v0 = ac[i], v1 = ac[i + stride], v2 = ac[i + 2 * stride], v3 = ac[i + 3 * stride];
            oc[i] = __dmul_rn(v0, bj); oc[i + stride] = __dmul_rn(v1, bj);
            oc[i + 2 * stride] = __dmul_rn(v2, bj); oc[i + 3 * stride] = __dmul_rn(v3, bj);
        }
        for (; i < ni; i += stride) oc[i] = __dmul_rn(ac[i], bj);
    }
}

extern "C" int rb_einsum_ij_j(rb_ctx *ctx, const double *a, int64_t lda, const double *b, double *out, int64_t ldo,
                              int64_t ni, int64_t nj)
{
    RB_REQUIRE(ctx && ni >= 0 && nj >= 0, "rb_einsum_ij_j: bad arguments");
    if (ni == 0 || nj == 0) return RB_OK;
    RB_REQUIRE(a && b && out && lda >= ni && ldo >= ni, "rb_einsum_ij_j: NULL buffer or leading dimension too small");
    RB_CUDA(cudaSetDevice(ctx->device));
    i64 bx = rb_cdiv(ni, 1024);
    if (bx > 64) bx = 64;
    i64 by = nj < 65535 ? nj : 65535;
    const i64 cap = (i64)ctx->num_sms * 8;
    if (bx * by > cap) by = cap / bx > 0 ? cap / bx : 1;
    rb_einsum_ij_j_kernel<<<dim3((unsigned)bx, (unsigned)by), 256, 0, ctx->stream>>>(a, lda, b, out, ldo, ni, nj);
    RB_LAUNCHED(ctx);
    return RB_OK;
}

__global__ void __launch_bounds__(256) rb_einsum_i_j_kernel(const double *__restrict__ a, const double *__restrict__ b,
                                                            double *__restrict__ out, i64 ni, i64 nj)
{
    for (i64 j = blockIdx.y; j < nj; j += gridDim.y) {
        const double bj = b[j];
        double *oc = out + j * ni;
        for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < ni; i += (i64)gridDim.x * blockDim.x)
            oc[i] = __dmul_rn(a[i], bj);
    }
}

extern "C" int rb_einsum_i_j(rb_ctx *ctx, const double *a, const double *b, double *out, int64_t ni, int64_t nj)
{
    RB_REQUIRE(ctx && ni >= 0 && nj >= 0, "rb_einsum_i_j: bad arguments");
    if (ni == 0 || nj == 0) return RB_OK;
    RB_REQUIRE(a && b && out, "rb_einsum_i_j: NULL buffer");
    RB_CUDA(cudaSetDevice(ctx->device));
    i64 bx = rb_cdiv(ni, 256);
    if (bx > 64) bx = 64;
    i64 by = nj < 65535 ? nj : 65535;
    const i64 cap = (i64)ctx->num_sms * 8;
    if (bx * by > cap) by = cap / bx > 0 ? cap / bx : 1;
    rb_einsum_i_j_kernel<<<dim3((unsigned)bx, (unsigned)by), 256, 0, ctx->stream>>>(a, b, out, ni, nj);
    RB_LAUNCHED(ctx);
    return RB_OK;
}

// grid (chunks, columns): CTA reduces rows [chunk*rows_per_chunk, ...) of column p of both operands.
__global__ void __launch_bounds__(256) rb_einsum_ip_ip_kernel(const double *__restrict__ a, i64 lda,
                                                              const double *__restrict__ b, i64 ldb, i64 ni,
                                                              i64 rows_per_chunk, double *__restrict__ partial, i64 chunks)
{
    const i64 p = blockIdx.y, chunk = blockIdx.x;
    const i64 r0 = chunk * rows_per_chunk;
    i64 r1 = r0 + rows_per_chunk;
    if (r1 > ni) r1 = ni;
    const double *ac = a + p * lda, *bc = b + p * ldb;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    i64 i = r0 + threadIdx.x;
    for (; i + 768 < r1; i += 1024) {
        const double a0 = ac[i], a1 = ac[i + 256], a2 = ac[i + 512], a3 = ac[i + 768];
        const double b0 = bc[i], b1 = bc[i + 256], b2 = bc[i + 512], b3 = bc[i + 768];
        s0 += a0 * b0; s1 += a1 * b1; s2 += a2 * b2; s3 += a3 * b3;
    }
    for (; i < r1; i += 256) s0 += ac[i] * bc[i];
    double s = (s0 + s1) + (s2 + s3);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __shared__ double red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[w];
        partial[chunk + p * chunks] = t;
    }
}

__global__ void __launch_bounds__(256) rb_einsum_ip_ip_finish_kernel(const double *__restrict__ partial, i64 chunks, i64 np,
                                                                     double *__restrict__ out)
{
    for (i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x; p < np; p += (i64)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (i64 c = 0; c < chunks; ++c) s += partial[c + p * chunks]; // fixed order => deterministic
        out[p] = s;
    }
}

extern "C" int rb_einsum_ip_ip(rb_ctx *ctx, const double *a, int64_t lda, const double *b, int64_t ldb, double *out,
                               int64_t ni, int64_t np)
{
    RB_REQUIRE(ctx && ni >= 0 && np >= 0, "rb_einsum_ip_ip: bad arguments");
    if (np == 0) return RB_OK;
    RB_REQUIRE(out, "rb_einsum_ip_ip: out is NULL");
    RB_CUDA(cudaSetDevice(ctx->device));
    if (ni == 0) return rb_scale_or_zero(ctx, out, np, 1, 0.0);
    RB_REQUIRE(a && b && lda >= ni && ldb >= ni, "rb_einsum_ip_ip: NULL buffer or leading dimension too small");
    // enough CTAs to fill the chip: split long columns into row chunks (multiples of 1024 rows)
    i64 chunks = 1;
    const i64 want = (i64)ctx->num_sms * 8;
    if (np < want) chunks = rb_cdiv(want, np);
    i64 rows_per_chunk = rb_cdiv(rb_cdiv(ni, chunks), 1024) * 1024;
    chunks = rb_cdiv(ni, rows_per_chunk);
    RB_REQUIRE(np <= 65535 * (i64)32768, "rb_einsum_ip_ip: too many columns");
    void *ws;
    RB_TRY(rb_ws_reserve(ctx, 1, chunks * np * 8, &ws));
    for (i64 p0 = 0; p0 < np; p0 += 65535) {
        const i64 pn = np - p0 < 65535 ? np - p0 : 65535;
        rb_einsum_ip_ip_kernel<<<dim3((unsigned)chunks, (unsigned)pn), 256, 0, ctx->stream>>>(
            a + p0 * lda, lda, b + p0 * ldb, ldb, ni, rows_per_chunk, (double *)ws + p0 * chunks, chunks);
        RB_LAUNCHED(ctx);
    }
    i64 fb = rb_cdiv(np, 256);
    if (fb > (i64)ctx->num_sms * 8) fb = (i64)ctx->num_sms * 8;
    rb_einsum_ip_ip_finish_kernel<<<(unsigned)fb, 256, 0, ctx->stream>>>((const double *)ws, chunks, np, out);
    RB_LAUNCHED(ctx);
    return RB_OK;
}
