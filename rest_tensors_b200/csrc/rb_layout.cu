// rb_layout.cu -- the HBM-bound, bit-exact half of the hot path:
//   MatrixUpper pack/unpack, per-slab symmetric pack, strided sub-box copies (copy_mm/mr/rm/rr),
//   rank-3 transposes, the axpy family and the counter-based synthetic generators.
// Every kernel moves each algorithmic byte once, coalesced along the unit-stride index.
#include "rb_common.cuh"

// ---------------------------------------------------------------------------------------------------------
// pack: packed[j(j+1)/2 + i] = full[i + j*n], i <= j      (matrixfull.rs:638-646, matrix_trait.rs:191-217)
// One CTA per (64-row block, column); tiles strictly below the diagonal exit at once.  blockIdx.z = slab.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64) rb_pack_upper_kernel(const double *__restrict__ full, double *__restrict__ packed,
                                                           i64 n, i64 full_slab, i64 packed_slab)
{
    i64 j = blockIdx.y;
    i64 i = (i64)blockIdx.x * 64 + threadIdx.x;
    if (i > j) return;
    const double *f = full + (i64)blockIdx.z * full_slab;
    double *p = packed + (i64)blockIdx.z * packed_slab;
    p[j * (j + 1) / 2 + i] = f[i + j * n];
}

// Wider variant for large n: each CTA handles a 256-row x 8-column panel (fewer, fatter CTAs).
__global__ void __launch_bounds__(256) rb_pack_upper_panel_kernel(const double *__restrict__ full,
                                                                  double *__restrict__ packed, i64 n, i64 full_slab,
                                                                  i64 packed_slab)
{
    i64 j0 = (i64)blockIdx.y * 8;
    i64 i = (i64)blockIdx.x * 256 + threadIdx.x;
    if ((i64)blockIdx.x * 256 > j0 + 7) return;
    const double *f = full + (i64)blockIdx.z * full_slab;
    double *p = packed + (i64)blockIdx.z * packed_slab;
    double v[8];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
        i64 j = j0 + jj;
        v[jj] = (j < n && i <= j) ? f[i + j * n] : 0.0;
    }
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
        i64 j = j0 + jj;
        if (j < n && i <= j) p[j * (j + 1) / 2 + i] = v[jj];
    }
}


static int launch_pack(rb_ctx *ctx, const double *full, i64 n, i64 nslab, double *packed)
{
    if (n == 0 || nslab == 0) return RB_OK;
    i64 np = n * (n + 1) / 2;
    // (Two 256-bit variants were measured and dropped, profiles/r02_hbm_kernels.md: 32-byte loads with each thread storing its own
    //  4 elements -- 16 scattered sectors per store instruction -- and 32-byte loads staged through shared memory with one 512-byte
    //  run per store instruction reached 5.3 / 4.9 TB/s at n = 8000 against 5.9 TB/s for the kernels below, whose 8-byte accesses
    //  are fully coalesced on both sides; the packed runs start at arbitrary 8-byte offsets, so wider stores buy nothing.)
    for (i64 z0 = 0; z0 < nslab; z0 += 65535) {
        unsigned nz = (unsigned)((nslab - z0) < 65535 ? (nslab - z0) : 65535);
        if (n >= 512) {
            dim3 grid((unsigned)rb_cdiv(n, 256), (unsigned)rb_cdiv(n, 8), nz);
            RB_REQUIRE(grid.y <= 65535, "pack: n too large");
            rb_pack_upper_panel_kernel<<<grid, 256, 0, ctx->stream>>>(full + z0 * n * n, packed + z0 * np, n, n * n, np);
        } else {
            dim3 grid((unsigned)rb_cdiv(n, 64), (unsigned)n, nz);
            rb_pack_upper_kernel<<<grid, 64, 0, ctx->stream>>>(full + z0 * n * n, packed + z0 * np, n, n * n, np);
        }
        RB_LAUNCHED(ctx);
    }
    return RB_OK;
}

extern "C" int rb_pack_upper(rb_ctx *ctx, const double *full, int64_t n, double *packed)
{
    RB_REQUIRE(ctx && n >= 0, "rb_pack_upper: bad arguments");
    RB_REQUIRE(n == 0 || (full && packed), "rb_pack_upper: NULL buffer");
    RB_CUDA(cudaSetDevice(ctx->device));
    return launch_pack(ctx, full, n, 1, packed);
}

extern "C" int rb_ri_pack_symm(rb_ctx *ctx, const double *ri, int64_t nao, int64_t naux, double *out)
{
    RB_REQUIRE(ctx && nao >= 0 && naux >= 0, "rb_ri_pack_symm: bad arguments");
    RB_CUDA(cudaSetDevice(ctx->device));
    return launch_pack(ctx, ri, nao, naux, out);
}

// ---------------------------------------------------------------------------------------------------------
// unpack: full[i + j*n] = full[j + i*n] = packed[j(j+1)/2 + i], i <= j          (matrixupper.rs:330-373)
// One CTA per 64x64 tile pair (ti <= tj): the packed tile is read once (coalesced along i, 16 loads in flight per
// thread), parked in shared memory, written to the upper position and -- transposed -- to the mirrored lower
// position, both with 16-byte stores when n is even.  The reference's mirror loop reads stride-n; this reads each
// packed byte once and writes each full byte once.
// ---------------------------------------------------------------------------------------------------------
constexpr int UT = 64; // tile edge

template <bool VEC>
__global__ void __launch_bounds__(256) rb_unpack_upper_kernel(const double *__restrict__ packed,
                                                              double *__restrict__ full, i64 n)
{
    __shared__ double tile[UT][UT + 1]; // [column jj][row ii]
    // linear tile-pair index -> (ti <= tj): pair p = tj(tj+1)/2 + ti
    const i64 p = blockIdx.x;
    i64 tj = (i64)((sqrt(8.0 * (double)p + 1.0) - 1.0) * 0.5);
    while (tj * (tj + 1) / 2 > p) --tj;
    while ((tj + 1) * (tj + 2) / 2 <= p) ++tj;
    const i64 ti = p - tj * (tj + 1) / 2;
    const i64 i0 = ti * UT, j0 = tj * UT;
    {
        const int tx = threadIdx.x & (UT - 1), ty = threadIdx.x >> 6; // 64 rows x 4 columns per pass
        const i64 i = i0 + tx;
        double v[UT / 4];
#pragma unroll
        for (int r = 0; r < UT / 4; ++r) {
            const i64 j = j0 + ty + r * 4;
            v[r] = 0.0;
            if (i < n && j < n) {
                // off-diagonal tiles have i < j everywhere; the diagonal tile fetches the (min, max) element
                const i64 lo = i <= j ? i : j, hi = i <= j ? j : i;
                v[r] = packed[hi * (hi + 1) / 2 + lo];
            }
        }
#pragma unroll
        for (int r = 0; r < UT / 4; ++r) tile[ty + r * 4][tx] = v[r];
    }
    __syncthreads();
    if (VEC) { // n even: rows (2t, 2t+1) of every column start 16-byte aligned
        const int t2 = threadIdx.x & 31, ty = threadIdx.x >> 5; // 32 row pairs x 8 columns per pass
#pragma unroll
        for (int r = 0; r < UT / 8; ++r) {
            const int jj = ty + r * 8;
            const i64 i = i0 + 2 * t2, j = j0 + jj;
            if (i < n && j < n) // n even and i even => i + 1 < n as well
                *reinterpret_cast<double2 *>(full + i + j * n) = make_double2(tile[jj][2 * t2], tile[jj][2 * t2 + 1]);
        }
        if (ti == tj) return; // the diagonal tile is already symmetric
#pragma unroll
        for (int r = 0; r < UT / 8; ++r) {
            const int ii = ty + r * 8; // column of the mirrored tile = row of the upper tile
            const i64 row = j0 + 2 * t2, col = i0 + ii;
            if (row < n && col < n)
                *reinterpret_cast<double2 *>(full + row + col * n) = make_double2(tile[2 * t2][ii], tile[2 * t2 + 1][ii]);
        }
    } else {
        const int tx = threadIdx.x & (UT - 1), ty = threadIdx.x >> 6;
#pragma unroll
        for (int r = 0; r < UT / 4; ++r) {
            const int jj = ty + r * 4;
            const i64 i = i0 + tx, j = j0 + jj;
            if (i < n && j < n) full[i + j * n] = tile[jj][tx];
        }
        if (ti == tj) return;
#pragma unroll
        for (int r = 0; r < UT / 4; ++r) {
            const int ii = ty + r * 4;
            const i64 row = j0 + tx, col = i0 + ii;
            if (row < n && col < n) full[row + col * n] = tile[tx][ii];
        }
    }
}


// (Occupancy: these kernels run 4 CTAs per SM -- 58-62 registers, 32 KB of exchange buffer.  Compiling them for 5 or 6 CTAs per SM
//  (`__launch_bounds__(256, 5 / 6)`: 48 / 40 registers, a few spilled bytes) with the maximum shared-memory carve-out was measured at
//  4.1-4.9 TB/s against 5.8-6.3 (tools/gpu_batch_r02ac.sh): the smaller L1 hurts the 32-byte streams more than the extra CTAs help.)
// ---- 32-byte exchange through shared memory (shared by the 256-bit unpack and transpose kernels) ---------------------
// A 64 x 64 tile is 16 x 16 micro-tiles of 4 x 4; thread (lr, lc) = (tid & 15, tid >> 4) owns micro-tile rows 4 lr .. 4 lr + 3,
// columns 4 lc .. 4 lc + 3.  After the register transpose it holds four 32-byte units u[e] = (column quad lc of TRANSPOSED
// row R = 4 lr + e).  They go to slot [R][lc ^ lr] (and the two 16-byte halves of a unit swap places when bit 2 of lr and lc
// differ), which makes the writes (16 lanes: same lc, all lr) and the reads (16 lanes: same R, all lc) both conflict-free.
__device__ __forceinline__ void rb_xchg_write(double *sm, int lr, int lc, int e, const rb_d4 &u)
{
    const int R = 4 * lr + e, col = lc ^ lr, swap = ((lr >> 2) ^ (lc >> 2)) & 1;
    double *slot = sm + (R * 16 + col) * 4;
    *reinterpret_cast<double2 *>(slot + 2 * swap) = make_double2(u.x, u.y);
    *reinterpret_cast<double2 *>(slot + 2 * (1 - swap)) = make_double2(u.z, u.w);
}
__device__ __forceinline__ rb_d4 rb_xchg_read(const double *sm, int R, int lc)
{
    const int lr = (R >> 2) & 15, col = lc ^ lr, swap = ((lr >> 2) ^ (lc >> 2)) & 1;
    const double *slot = sm + (R * 16 + col) * 4;
    const double2 a = *reinterpret_cast<const double2 *>(slot + 2 * swap);
    const double2 b = *reinterpret_cast<const double2 *>(slot + 2 * (1 - swap));
    rb_d4 u; u.x = a.x; u.y = a.y; u.z = b.x; u.w = b.y;
    return u;
}

// four consecutive packed elements starting at an 8-byte aligned address: 16-byte loads where the parity allows
__device__ __forceinline__ rb_d4 rb_load_run4(const double *src)
{
    rb_d4 v;
    if ((((uintptr_t)src) & 15) == 0) {
        const double2 a = *reinterpret_cast<const double2 *>(src), b = *reinterpret_cast<const double2 *>(src + 2);
        v.x = a.x; v.y = a.y; v.z = b.x; v.w = b.y;
    } else {
        const double2 m = *reinterpret_cast<const double2 *>(src + 1);
        v.x = src[0]; v.y = m.x; v.z = m.y; v.w = src[3];
    }
    return v;
}

// 256-bit unpack (n % 4 == 0, `full` 32-byte aligned): one CTA per 64 x 64 tile pair (ti <= tj).  A thread reads the packed
// 4 x 4 micro-tile once, stores it to the upper position with four 32-byte stores (coalesced along i), transposes it in
// registers and hands the 32-byte units through shared memory to the lanes that store the mirrored tile, coalesced along j.
__global__ void __launch_bounds__(256) rb_unpack_upper4_kernel(const double *__restrict__ packed, double *__restrict__ full, i64 n)
{
    __shared__ __align__(32) double sm[64 * 64];
    const i64 p = blockIdx.x;
    i64 tj = (i64)((sqrt(8.0 * (double)p + 1.0) - 1.0) * 0.5);
    while (tj * (tj + 1) / 2 > p) --tj;
    while ((tj + 1) * (tj + 2) / 2 <= p) ++tj;
    const i64 ti = p - tj * (tj + 1) / 2;
    const bool diag = ti == tj;
    const int lr = threadIdx.x & 15, lc = threadIdx.x >> 4;
    const i64 i = ti * 64 + 4 * lr, j0 = tj * 64 + 4 * lc;
    // diagonal tile: micro-tiles below the diagonal (lr > lc) are produced by the mirror of (lc, lr)
    const bool mine = i < n && j0 < n && (!diag || lr <= lc);
    rb_d4 v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) { v[e].x = 0.0; v[e].y = 0.0; v[e].z = 0.0; v[e].w = 0.0; }
    if (mine) {
        if (!diag || lr < lc) {
#pragma unroll
            for (int e = 0; e < 4; ++e) { const i64 j = j0 + e; v[e] = rb_load_run4(packed + j * (j + 1) / 2 + i); }
        } else { // the 4 x 4 block on the diagonal: rows i .. j of column j, mirrored inside the registers
            double m[4][4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const i64 j = j0 + e;
#pragma unroll
                for (int k = 0; k < 4; ++k) m[e][k] = (k <= e) ? packed[j * (j + 1) / 2 + i + k] : 0.0;
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                v[e].x = m[e][0];
                v[e].y = e >= 1 ? m[e][1] : m[1][e];
                v[e].z = e >= 2 ? m[e][2] : m[2][e];
                v[e].w = e >= 3 ? m[e][3] : m[3][e];
            }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) rb_st256(full + i + (j0 + e) * n, v[e]);
    }
    // mirrored tile: full[j0 .. j0+3 + (i + k) * n] = (v[0].k, v[1].k, v[2].k, v[3].k)
    const bool give = mine && (!diag || lr < lc);
    if (give) {
        rb_d4 u;
        u.x = v[0].x; u.y = v[1].x; u.z = v[2].x; u.w = v[3].x; rb_xchg_write(sm, lr, lc, 0, u);
        u.x = v[0].y; u.y = v[1].y; u.z = v[2].y; u.w = v[3].y; rb_xchg_write(sm, lr, lc, 1, u);
        u.x = v[0].z; u.y = v[1].z; u.z = v[2].z; u.w = v[3].z; rb_xchg_write(sm, lr, lc, 2, u);
        u.x = v[0].w; u.y = v[1].w; u.z = v[2].w; u.w = v[3].w; rb_xchg_write(sm, lr, lc, 3, u);
    }
    __syncthreads();
    const int qc = threadIdx.x & 15; // column quad (along j) this lane stores
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
        const int R = (threadIdx.x >> 4) + 16 * pass;   // row of the upper tile = column of the mirrored one
        const i64 col = ti * 64 + R, row = tj * 64 + 4 * qc;
        if (col >= n || row >= n) continue;
        if (diag && !((R >> 2) < qc)) continue;        // only the slots strictly-upper micro-tiles wrote
        rb_st256(full + row + col * n, rb_xchg_read(sm, R, qc));
    }
}

extern "C" int rb_unpack_upper(rb_ctx *ctx, const double *packed, int64_t n, double *full)
{
    RB_REQUIRE(ctx && n >= 0, "rb_unpack_upper: bad arguments");
    if (n == 0) return RB_OK;
    RB_REQUIRE(packed && full, "rb_unpack_upper: NULL buffer");
    RB_CUDA(cudaSetDevice(ctx->device));
    i64 nt = rb_cdiv(n, UT);
    i64 pairs = nt * (nt + 1) / 2;
    RB_REQUIRE(pairs < 2147483647LL, "rb_unpack_upper: n too large");
    if ((n & 3) == 0 && rb_aligned32(full) && n >= 64)
        rb_unpack_upper4_kernel<<<(unsigned)pairs, 256, 0, ctx->stream>>>(packed, full, n);
    else if ((n & 1) == 0 && (((uintptr_t)full) & 15) == 0)
        rb_unpack_upper_kernel<true><<<(unsigned)pairs, 256, 0, ctx->stream>>>(packed, full, n);
    else
        rb_unpack_upper_kernel<false><<<(unsigned)pairs, 256, 0, ctx->stream>>>(packed, full, n);
    RB_LAUNCHED(ctx);
    return RB_OK;
}

// Copy one triangle of a square matrix onto the other (used after SYRK-style K builds).
__global__ void __launch_bounds__(256) rb_symmetrize_kernel(double *__restrict__ c, i64 n, i64 ldc, int from_upper)
{
    __shared__ double tile[32][33];
    i64 p = blockIdx.x;
    i64 tj = (i64)((sqrt(8.0 * (double)p + 1.0) - 1.0) * 0.5);
    while (tj * (tj + 1) / 2 > p) --tj;
    while ((tj + 1) * (tj + 2) / 2 <= p) ++tj;
    i64 ti = p - tj * (tj + 1) / 2; // ti <= tj
    int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    // source tile: upper => rows from block ti, cols from block tj ; lower => rows tj, cols ti
    i64 sr0 = from_upper ? ti * 32 : tj * 32, sc0 = from_upper ? tj * 32 : ti * 32;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        int cc = ty + r * 8;
        i64 row = sr0 + tx, col = sc0 + cc;
        // only the source triangle is read: the other half of a diagonal tile may never have been written (SYRK with beta == 0)
        const bool in_src = from_upper ? (row <= col) : (row >= col);
        tile[cc][tx] = (row < n && col < n && in_src) ? c[row + col * ldc] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        int cc = ty + r * 8;
        i64 row = sc0 + tx, col = sr0 + cc; // destination (transposed position)
        if (row < n && col < n) {
            bool dst_is_target = from_upper ? (row > col) : (row < col);
            if (dst_is_target) c[row + col * ldc] = tile[tx][cc];
        }
    }
}

int rb_symmetrize(rb_ctx *ctx, double *c, i64 n, i64 ldc, bool from_upper)
{
    if (n <= 1) return RB_OK;
    i64 nt = rb_cdiv(n, 32);
    i64 pairs = nt * (nt + 1) / 2;
    rb_symmetrize_kernel<<<(unsigned)pairs, 256, 0, ctx->stream>>>(c, n, ldc, from_upper ? 1 : 0);
    RB_LAUNCHED(ctx);
    return RB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Generic strided 3-D copy -- copy_mm / copy_mr / copy_rm / copy_rr and transpose_ikj all reduce to it.
// A CTA is a (lanes along i) x (rows) thread grid; rows (j, k) are decoded once per row, never per element, and a
// thread keeps 4 independent loads in flight.  VEC=2 moves double2 along i when both sides are unit-stride and
// 16-byte aligned.
// ---------------------------------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(256) rb_copy3d_kernel(const double *__restrict__ src, i64 si, i64 sj, i64 sk,
                                                        double *__restrict__ dst, i64 di, i64 dj, i64 dk, i64 ni,
                                                        i64 nj, i64 nk, int lg_lanes)
{
    const int lanes = 1 << lg_lanes;                 // threads along i (power of two <= 256)
    const int lane = threadIdx.x & (lanes - 1);
    const int rsub = threadIdx.x >> lg_lanes;        // row handled by this thread inside the CTA's row group
    const int rows_per_cta = 256 >> lg_lanes;
    const i64 niv = ni / VEC;
    const i64 rows = nj * nk;
    for (i64 r = (i64)blockIdx.x * rows_per_cta + rsub; r < rows; r += (i64)gridDim.x * rows_per_cta) {
        const i64 j = r % nj, k = r / nj;
        const double *s = src + j * sj + k * sk;
        double *d = dst + j * dj + k * dk;
        i64 i = lane;
        if (VEC == 4) {
            for (; i + 3 * lanes < niv; i += 4 * lanes) {
                const rb_d4 v0 = rb_ld256(s + 4 * i), v1 = rb_ld256(s + 4 * (i + lanes)), v2 = rb_ld256(s + 4 * (i + 2 * lanes)),
                            v3 = rb_ld256(s + 4 * (i + 3 * lanes));
                rb_st256(d + 4 * i, v0); rb_st256(d + 4 * (i + lanes), v1); rb_st256(d + 4 * (i + 2 * lanes), v2);
                rb_st256(d + 4 * (i + 3 * lanes), v3);
            }
            for (; i < niv; i += lanes) rb_st256(d + 4 * i, rb_ld256(s + 4 * i));
        } else if (VEC == 2) {
            const double2 *s2 = reinterpret_cast<const double2 *>(s);
            double2 *d2 = reinterpret_cast<double2 *>(d);
            for (; i + 3 * lanes < niv; i += 4 * lanes) {
                const double2 v0 = s2[i], v1 = s2[i + lanes], v2 = s2[i + 2 * lanes], v3 = s2[i + 3 * lanes];
                d2[i] = v0; d2[i + lanes] = v1; d2[i + 2 * lanes] = v2; d2[i + 3 * lanes] = v3;
            }
            for (; i < niv; i += lanes) d2[i] = s2[i];
        } else {
            for (; i + 3 * lanes < niv; i += 4 * lanes) {
                const double v0 = s[i * si], v1 = s[(i + lanes) * si], v2 = s[(i + 2 * lanes) * si], v3 = s[(i + 3 * lanes) * si];
                d[i * di] = v0; d[(i + lanes) * di] = v1; d[(i + 2 * lanes) * di] = v2; d[(i + 3 * lanes) * di] = v3;
            }
            for (; i < niv; i += lanes) d[i * di] = s[i * si];
        }
    }
}

// one contiguous run of niv 32-byte vectors: grid-stride with 8 loads in flight per thread, each a whole grid apart -- the
// access pattern of the fastest variant of tools/micro/copy_bench.cu (6.6-6.7 TB/s with >= 16 CTAs per SM in the grid; taking
// contiguous chunks per CTA instead is no faster)
__global__ void __launch_bounds__(256) rb_copy_flat4_kernel(const double *__restrict__ s, double *__restrict__ d, i64 niv)
{
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < niv; i += 8 * stride) {
        rb_d4 v[8]; // the last pass is predicated, not serialised: its loads are in flight together as well
#pragma unroll
        for (int u = 0; u < 8; ++u) if (i + u * stride < niv) v[u] = rb_ld256(s + 4 * (i + u * stride));
#pragma unroll
        for (int u = 0; u < 8; ++u) if (i + u * stride < niv) rb_st256(d + 4 * (i + u * stride), v[u]);
    }
}

int rb_copy3d(rb_ctx *ctx, const double *src, i64 s0, i64 si, i64 sj, i64 sk, double *dst, i64 d0, i64 di, i64 dj,
              i64 dk, i64 ni, i64 nj, i64 nk)
{
    if (ni <= 0 || nj <= 0 || nk <= 0) return RB_OK;
    const double *s = src + s0;
    double *d = dst + d0;
    if (si == 1 && di == 1) {
        // rows that follow each other without a gap on both sides are one longer row (whole slabs, whole tensors)
        if (sj == ni && dj == ni && nj > 1) { ni *= nj; nj = nk; sj = sk; dj = dk; nk = 1; sk = 0; dk = 0; }
        if (sj == ni && dj == ni && nj > 1) { ni *= nj; nj = 1; }
        // unit-stride runs on both sides: bulk-tensor copy when requested and TMA can describe the operands
        const int st = rb_tma_copy3d(ctx, s, sj, sk, d, dj, dk, ni, nj, nk);
        if (st != RB_TMA_NOT_ELIGIBLE) return st;
    }
    const bool vec4 = si == 1 && di == 1 && (ni % 4 == 0) && (sj % 4 == 0) && (sk % 4 == 0) && (dj % 4 == 0) && (dk % 4 == 0) &&
                      rb_aligned32(s) && rb_aligned32(d);
    if (vec4) {
        const i64 niv = ni / 4;
        int lg = 0;
        while (lg < 8 && ((i64)2 << lg) * 4 <= niv) ++lg;
        if (lg < 3 && niv >= 8) lg = 3;
        const i64 rows = nj * nk, rows_per_cta = 256 >> lg;
        // one row: split it over the CTAs (each takes whole 4-load passes of 256 lanes)
        if (rows == 1 && niv >= 4096) {
            i64 blocks = rb_cdiv(niv, 256 * 8);
            const i64 cap = (i64)ctx->num_sms * 32;
            if (blocks > cap) blocks = cap;
            rb_copy_flat4_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(s, d, niv);
            RB_LAUNCHED(ctx);
            return RB_OK;
        }
        i64 blocks = rb_cdiv(rows, rows_per_cta);
        const i64 cap = (i64)ctx->num_sms * 32;
        if (blocks > cap) blocks = cap;
        rb_copy3d_kernel<4><<<(unsigned)blocks, 256, 0, ctx->stream>>>(s, si, sj, sk, d, di, dj, dk, ni, nj, nk, lg);
        RB_LAUNCHED(ctx);
        return RB_OK;
    }
    bool vec = si == 1 && di == 1 && (ni % 2 == 0) && (sj % 2 == 0) && (sk % 2 == 0) && (dj % 2 == 0) &&
               (dk % 2 == 0) && (((uintptr_t)s & 15) == 0) && (((uintptr_t)d & 15) == 0);
    const i64 niv = vec ? ni / 2 : ni;
    int lg = 0;
    while (lg < 8 && ((i64)2 << lg) * 4 <= niv) ++lg; // lanes = largest power of two <= niv / 4 (4 loads in flight per
    if (lg < 3 && niv >= 8) lg = 3;                   // thread), at most 256; at least a quarter-warp per row
    const i64 rows = nj * nk, rows_per_cta = 256 >> lg;
    i64 blocks = rb_cdiv(rows, rows_per_cta);
    i64 cap = (i64)ctx->num_sms * 8;
    if (blocks > cap) blocks = cap;
    if (vec) rb_copy3d_kernel<2><<<(unsigned)blocks, 256, 0, ctx->stream>>>(s, si, sj, sk, d, di, dj, dk, ni, nj, nk, lg);
    else rb_copy3d_kernel<1><<<(unsigned)blocks, 256, 0, ctx->stream>>>(s, si, sj, sk, d, di, dj, dk, ni, nj, nk, lg);
    RB_LAUNCHED(ctx);
    return RB_OK;
}

static bool box_ok(i64 start, i64 len, i64 dim) { return start >= 0 && len >= 0 && start + len <= dim; }

extern "C" int rb_copy_mm(rb_ctx *ctx, int xl, int yl, const double *f, int fx, int fy, int fxs, int fys, double *t,
                          int tx, int ty, int txs, int tys)
{
    RB_REQUIRE(ctx, "rb_copy_mm: ctx is NULL");
    RB_REQUIRE(box_ok(fxs, xl, fx) && box_ok(fys, yl, fy) && box_ok(txs, xl, tx) && box_ok(tys, yl, ty),
               "rb_copy_mm: block outside matrix");
    RB_CUDA(cudaSetDevice(ctx->device));
    return rb_copy3d(ctx, f, fxs + (i64)fys * fx, 1, fx, 0, t, txs + (i64)tys * tx, 1, tx, 0, xl, yl, 1);
}

// strides of the (x1, x2) block inside an RI tensor [X,Y,Z] for the three copy modes (restmatr.f90:227-237)
static int ri_mode_strides(int mod, i64 X, i64 Y, i64 Z, i64 s1, i64 s2, i64 x3, i64 l1, i64 l2, i64 *off, i64 *st1,
                           i64 *st2)
{
    if (mod == 0) {
        if (!(box_ok(s1, l1, X) && box_ok(s2, l2, Y) && x3 >= 0 && x3 < Z)) return 1;
        *off = s1 + s2 * X + x3 * X * Y; *st1 = 1; *st2 = X;
    } else if (mod == 1) {
        if (!(box_ok(s1, l1, X) && box_ok(s2, l2, Z) && x3 >= 0 && x3 < Y)) return 1;
        *off = s1 + x3 * X + s2 * X * Y; *st1 = 1; *st2 = X * Y;
    } else {
        if (!(box_ok(s1, l1, Y) && box_ok(s2, l2, Z) && x3 >= 0 && x3 < X)) return 1;
        *off = x3 + s1 * X + s2 * X * Y; *st1 = X; *st2 = X * Y;
    }
    return 0;
}

extern "C" int rb_copy_mr(rb_ctx *ctx, int xl, int yl, const double *f, int fx, int fy, int fxs, int fys, double *t,
                          int tx, int ty, int tz, int txs, int tys, int t3, int mod)
{
    RB_REQUIRE(ctx, "rb_copy_mr: ctx is NULL");
    if (mod < 0 || mod > 2) return RB_OK; // restmatr.f90:227-237: other mods are a no-op
    RB_REQUIRE(box_ok(fxs, xl, fx) && box_ok(fys, yl, fy), "rb_copy_mr: block outside matrix");
    i64 off, s1, s2;
    RB_REQUIRE(ri_mode_strides(mod, tx, ty, tz, txs, tys, t3, xl, yl, &off, &s1, &s2) == 0,
               "rb_copy_mr: block outside tensor");
    RB_CUDA(cudaSetDevice(ctx->device));
    return rb_copy3d(ctx, f, fxs + (i64)fys * fx, 1, fx, 0, t, off, s1, s2, 0, xl, yl, 1);
}

extern "C" int rb_copy_rm(rb_ctx *ctx, int xl, int yl, const double *f, int fx, int fy, int fz, int fxs, int fys,
                          int f3, int mod, double *t, int tx, int ty, int txs, int tys)
{
    RB_REQUIRE(ctx, "rb_copy_rm: ctx is NULL");
    if (mod < 0 || mod > 2) return RB_OK;
    RB_REQUIRE(box_ok(txs, xl, tx) && box_ok(tys, yl, ty), "rb_copy_rm: block outside matrix");
    i64 off, s1, s2;
    RB_REQUIRE(ri_mode_strides(mod, fx, fy, fz, fxs, fys, f3, xl, yl, &off, &s1, &s2) == 0,
               "rb_copy_rm: block outside tensor");
    RB_CUDA(cudaSetDevice(ctx->device));
    return rb_copy3d(ctx, f, off, s1, s2, 0, t, txs + (i64)tys * tx, 1, tx, 0, xl, yl, 1);
}

extern "C" int rb_copy_rr(rb_ctx *ctx, int xl, int yl, int zl, const double *f, int fx, int fy, int fz, int fxs,
                          int fys, int fzs, double *t, int tx, int ty, int tz, int txs, int tys, int tzs)
{
    RB_REQUIRE(ctx, "rb_copy_rr: ctx is NULL");
    RB_REQUIRE(box_ok(fxs, xl, fx) && box_ok(fys, yl, fy) && box_ok(fzs, zl, fz) && box_ok(txs, xl, tx) &&
                   box_ok(tys, yl, ty) && box_ok(tzs, zl, tz),
               "rb_copy_rr: box outside tensor");
    RB_CUDA(cudaSetDevice(ctx->device));
    i64 FX = fx, FY = fy, TX = tx, TY = ty;
    return rb_copy3d(ctx, f, fxs + fys * FX + fzs * FX * FY, 1, FX, FX * FY, t, txs + tys * TX + tzs * TX * TY, 1, TX,
                     TX * TY, xl, yl, zl);
}

// ---------------------------------------------------------------------------------------------------------
// Batched tiled 2-D transpose: out[c + r*ors + b*obs] = in[r + c*ics + b*ibs]   (r, c unit-stride on in / out)
// 64x64 tiles through padded shared memory; VEC: 16-byte loads along r and 16-byte stores along c (all strides even,
// bases 16-byte aligned), 16 loads in flight per thread.
// ---------------------------------------------------------------------------------------------------------
constexpr int TT = 64;

template <bool VEC>
__global__ void __launch_bounds__(256) rb_transpose_kernel(const double *__restrict__ in, i64 ics, i64 ibs,
                                                           double *__restrict__ out, i64 ors, i64 obs, i64 nr, i64 nc,
                                                           i64 tiles_r, i64 tiles_c, i64 total_tiles)
{
    __shared__ double tile[TT][TT + 1]; // [c][r]
    for (i64 t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const i64 tr = t % tiles_r, rest = t / tiles_r;
        const i64 tc = rest % tiles_c, b = rest / tiles_c;
        const double *ib = in + b * ibs;
        double *ob = out + b * obs;
        const i64 r0 = tr * TT, c0 = tc * TT;
        if (VEC) {
            const int t2 = threadIdx.x & 31, ty = threadIdx.x >> 5; // 32 pairs x 8 per pass
            double2 v[TT / 8];
#pragma unroll
            for (int q = 0; q < TT / 8; ++q) {
                const i64 r = r0 + 2 * t2, c = c0 + ty + q * 8;
                v[q] = make_double2(0.0, 0.0);
                if (r + 1 < nr && c < nc) v[q] = *reinterpret_cast<const double2 *>(ib + r + c * ics);
                else if (r < nr && c < nc) v[q].x = ib[r + c * ics];
            }
#pragma unroll
            for (int q = 0; q < TT / 8; ++q) { tile[ty + q * 8][2 * t2] = v[q].x; tile[ty + q * 8][2 * t2 + 1] = v[q].y; }
            __syncthreads();
#pragma unroll
            for (int q = 0; q < TT / 8; ++q) {
                const int rr = ty + q * 8;
                const i64 r = r0 + rr, c = c0 + 2 * t2;
                if (r < nr) {
                    if (c + 1 < nc) *reinterpret_cast<double2 *>(ob + c + r * ors) = make_double2(tile[2 * t2][rr], tile[2 * t2 + 1][rr]);
                    else if (c < nc) ob[c + r * ors] = tile[2 * t2][rr];
                }
            }
        } else {
            const int tx = threadIdx.x & (TT - 1), ty = threadIdx.x >> 6; // 64 x 4 per pass
            double v[TT / 4];
#pragma unroll
            for (int q = 0; q < TT / 4; ++q) {
                const i64 r = r0 + tx, c = c0 + ty + q * 4;
                v[q] = (r < nr && c < nc) ? ib[r + c * ics] : 0.0;
            }
#pragma unroll
            for (int q = 0; q < TT / 4; ++q) tile[ty + q * 4][tx] = v[q];
            __syncthreads();
#pragma unroll
            for (int q = 0; q < TT / 4; ++q) {
                const int rr = ty + q * 4;
                const i64 r = r0 + rr, c = c0 + tx;
                if (r < nr && c < nc) ob[c + r * ors] = tile[tx][rr];
            }
        }
        __syncthreads();
    }
}


// 256-bit form (all extents and strides multiples of 4, 32-byte aligned bases): a thread loads a 4 x 4 micro-tile with four
// 32-byte loads along r (16 lanes cover 64 consecutive r of one column), transposes it in registers, and the 32-byte units
// change hands through shared memory (rb_xchg_*) so that the four 32-byte stores along c are coalesced as well.
__global__ void __launch_bounds__(256) rb_transpose4_kernel(const double *__restrict__ in, i64 ics, i64 ibs, double *__restrict__ out,
                                                            i64 ors, i64 obs, i64 nr, i64 nc, i64 tiles_r, i64 tiles_c, i64 total_tiles)
{
    __shared__ __align__(32) double sm[64 * 64];
    const int lr = threadIdx.x & 15, lc = threadIdx.x >> 4;
    for (i64 t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const i64 tr = t % tiles_r, rest = t / tiles_r;
        const i64 tc = rest % tiles_c, b = rest / tiles_c;
        const double *ib = in + b * ibs;
        double *ob = out + b * obs;
        const i64 r = tr * 64 + 4 * lr, c = tc * 64 + 4 * lc;
        if (r < nr && c < nc) { // nr, nc are multiples of 4: a micro-tile is inside or outside as a whole
            const rb_d4 v0 = rb_ld256(ib + r + c * ics), v1 = rb_ld256(ib + r + (c + 1) * ics), v2 = rb_ld256(ib + r + (c + 2) * ics),
                        v3 = rb_ld256(ib + r + (c + 3) * ics);
            rb_d4 u;
            u.x = v0.x; u.y = v1.x; u.z = v2.x; u.w = v3.x; rb_xchg_write(sm, lr, lc, 0, u);
            u.x = v0.y; u.y = v1.y; u.z = v2.y; u.w = v3.y; rb_xchg_write(sm, lr, lc, 1, u);
            u.x = v0.z; u.y = v1.z; u.z = v2.z; u.w = v3.z; rb_xchg_write(sm, lr, lc, 2, u);
            u.x = v0.w; u.y = v1.w; u.z = v2.w; u.w = v3.w; rb_xchg_write(sm, lr, lc, 3, u);
        }
        __syncthreads();
        const int qc = threadIdx.x & 15;
#pragma unroll
        for (int pass = 0; pass < 4; ++pass) {
            const int R = (threadIdx.x >> 4) + 16 * pass;
            const i64 orow = tr * 64 + R, ocol = tc * 64 + 4 * qc; // out[ocol .. ocol+3 + orow * ors]
            if (orow < nr && ocol < nc) rb_st256(ob + ocol + orow * ors, rb_xchg_read(sm, R, qc));
        }
        __syncthreads();
    }
}

int rb_transpose_batched(rb_ctx *ctx, const double *in, i64 ics, i64 ibs, double *out, i64 ors, i64 obs, i64 nr,
                         i64 nc, i64 nbatch)
{
    if (nr <= 0 || nc <= 0 || nbatch <= 0) return RB_OK;
    {
        const int st = rb_tma_transpose(ctx, in, ics, ibs, out, ors, obs, nr, nc, nbatch);
        if (st != RB_TMA_NOT_ELIGIBLE) return st;
    }
    i64 tiles_r = rb_cdiv(nr, TT), tiles_c = rb_cdiv(nc, TT);
    i64 total = tiles_r * tiles_c * nbatch;
    if (((nr | nc | ics | ors) & 3) == 0 && (nbatch == 1 || ((ibs | obs) & 3) == 0) && rb_aligned32(in) && rb_aligned32(out)) {
        i64 blocks = total;
        const i64 cap4 = (i64)ctx->num_sms * 32;
        if (blocks > cap4) blocks = cap4;
        rb_transpose4_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(in, ics, ibs, out, ors, obs, nr, nc, tiles_r, tiles_c, total);
        RB_LAUNCHED(ctx);
        return RB_OK;
    }
    i64 blocks = total;
    i64 cap = (i64)ctx->num_sms * 16;
    if (blocks > cap) blocks = cap;
    const bool vec = ((ics | ibs | ors | obs) & 1) == 0 && ((((uintptr_t)in) | ((uintptr_t)out)) & 15) == 0;
    if (vec)
        rb_transpose_kernel<true><<<(unsigned)blocks, 256, 0, ctx->stream>>>(in, ics, ibs, out, ors, obs, nr, nc, tiles_r,
                                                                            tiles_c, total);
    else
        rb_transpose_kernel<false><<<(unsigned)blocks, 256, 0, ctx->stream>>>(in, ics, ibs, out, ors, obs, nr, nc, tiles_r,
                                                                             tiles_c, total);
    RB_LAUNCHED(ctx);
    return RB_OK;
}

extern "C" int rb_matrix_transpose(rb_ctx *ctx, const double *in, int64_t rows, int64_t cols, double *out)
{
    RB_REQUIRE(ctx && rows >= 0 && cols >= 0, "rb_matrix_transpose: bad arguments");
    RB_CUDA(cudaSetDevice(ctx->device));
    return rb_transpose_batched(ctx, in, rows, 0, out, cols, 0, rows, cols, 1);
}

// ri.rs:227-294.  in [I,J,K]:  0 jik -> [J,I,K]; 1 jki -> [J,K,I]; 2 kji -> [K,J,I]; 3 ikj -> [I,K,J]
extern "C" int rb_ri_transpose(rb_ctx *ctx, const double *in, int64_t I, int64_t J, int64_t K, int which, double *out)
{
    RB_REQUIRE(ctx && I >= 0 && J >= 0 && K >= 0, "rb_ri_transpose: bad arguments");
    RB_REQUIRE(which >= 0 && which <= 3, "rb_ri_transpose: which must be 0..3");
    RB_CUDA(cudaSetDevice(ctx->device));
    switch (which) {
    case 0: // per-slab transpose: out[j + i*J + k*IJ]
        return rb_transpose_batched(ctx, in, I, I * J, out, J, I * J, I, J, K);
    case 1: // [I,(JK)] -> [(JK),I]: out[n + i*JK], n = j + k*J
        return rb_transpose_batched(ctx, in, I, 0, out, J * K, 0, I, J * K, 1);
    case 2: // for each j: (i,k) -> (k,i): in i + k*IJ (+ j*I), out k + i*JK (+ j*K)
        return rb_transpose_batched(ctx, in, I * J, I, out, J * K, K, I, K, J);
    default: // ikj: out[i + k*I + j*IK] = in[i + j*I + k*IJ]; unit-stride runs of I
        return rb_copy3d(ctx, in, 0, 1, I, I * J, out, 0, 1, I * K, I, I, J, K);
    }
}

// ---------------------------------------------------------------------------------------------------------
// axpy family.  __dmul_rn/__dadd_rn keep the reference's unfused "*c += p*b" rounding (Rust never emits FMA).
// ---------------------------------------------------------------------------------------------------------
template <int OP>
__global__ void __launch_bounds__(256) rb_axpy_kernel(double *__restrict__ c, const double *__restrict__ p, double a,
                                                      double b, i64 n)
{
    i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double cv = c[i];
        double pv = (OP == 2) ? 0.0 : p[i];
        double r;
        if (OP == 0) r = __dadd_rn(cv, __dmul_rn(pv, b));                       // c += p*b
        else if (OP == 1) r = __dadd_rn(__dmul_rn(cv, a), __dmul_rn(pv, b));    // c = c*a + p*b
        else if (OP == 2) r = __dmul_rn(cv, a);                                 // c *= a
        else if (OP == 3) r = __dadd_rn(cv, pv);                                // c += p
        else r = __dsub_rn(cv, pv);                                             // c -= p
        c[i] = r;
    }
}

template <int OP>
__device__ __forceinline__ double rb_axpy_op(double cv, double pv, double a, double b)
{
    if (OP == 0) return __dadd_rn(cv, __dmul_rn(pv, b));
    if (OP == 1) return __dadd_rn(__dmul_rn(cv, a), __dmul_rn(pv, b));
    if (OP == 2) return __dmul_rn(cv, a);
    if (OP == 3) return __dadd_rn(cv, pv);
    return __dsub_rn(cv, pv);
}

// 256-bit form: n4 vectors of 4 doubles (c, p 32-byte aligned), two vectors of each operand in flight per thread
template <int OP>
__global__ void __launch_bounds__(256) rb_axpy4_kernel(double *__restrict__ c, const double *__restrict__ p, double a, double b, i64 n4)
{
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += 2 * stride) {
        const bool two = i + stride < n4;
        rb_d4 c0 = rb_ld256(c + 4 * i), c1, p0, p1;
        if (OP != 2) p0 = rb_ld256(p + 4 * i);
        if (two) { c1 = rb_ld256(c + 4 * (i + stride)); if (OP != 2) p1 = rb_ld256(p + 4 * (i + stride)); }
        c0.x = rb_axpy_op<OP>(c0.x, p0.x, a, b); c0.y = rb_axpy_op<OP>(c0.y, p0.y, a, b);
        c0.z = rb_axpy_op<OP>(c0.z, p0.z, a, b); c0.w = rb_axpy_op<OP>(c0.w, p0.w, a, b);
        rb_st256(c + 4 * i, c0);
        if (two) {
            c1.x = rb_axpy_op<OP>(c1.x, p1.x, a, b); c1.y = rb_axpy_op<OP>(c1.y, p1.y, a, b);
            c1.z = rb_axpy_op<OP>(c1.z, p1.z, a, b); c1.w = rb_axpy_op<OP>(c1.w, p1.w, a, b);
            rb_st256(c + 4 * (i + stride), c1);
        }
    }
}

template <int OP>
static int launch_axpy(rb_ctx *ctx, double *c, const double *p, double a, double b, i64 n)
{
    RB_REQUIRE(ctx && n >= 0, "axpy: bad arguments");
    if (n == 0) return RB_OK;
    RB_CUDA(cudaSetDevice(ctx->device));
    if (n >= 4096 && rb_aligned32(c) && (OP == 2 || rb_aligned32(p))) { // 256-bit body, scalar tail (< 4 elements)
        const i64 n4 = n / 4;
        i64 blocks4 = rb_cdiv(n4, 512);
        const i64 cap4 = (i64)ctx->num_sms * 32;
        if (blocks4 > cap4) blocks4 = cap4;
        rb_axpy4_kernel<OP><<<(unsigned)blocks4, 256, 0, ctx->stream>>>(c, p, a, b, n4);
        RB_LAUNCHED(ctx);
        if (n4 * 4 == n) return RB_OK;
        c += n4 * 4; if (OP != 2) p += n4 * 4; n -= n4 * 4;
    }
    i64 blocks = rb_cdiv(n, 256);
    i64 cap = (i64)ctx->num_sms * 16;
    if (blocks > cap) blocks = cap;
    rb_axpy_kernel<OP><<<(unsigned)blocks, 256, 0, ctx->stream>>>(c, p, a, b, n);
    RB_LAUNCHED(ctx);
    return RB_OK;
}
extern "C" int rb_self_scaled_add(rb_ctx *ctx, double *c, const double *p, double b, int64_t n) { return launch_axpy<0>(ctx, c, p, 0.0, b, n); }
extern "C" int rb_self_general_add(rb_ctx *ctx, double *c, const double *p, double a, double b, int64_t n) { return launch_axpy<1>(ctx, c, p, a, b, n); }
extern "C" int rb_self_multiple(rb_ctx *ctx, double *c, double a, int64_t n) { return launch_axpy<2>(ctx, c, nullptr, a, 0.0, n); }
extern "C" int rb_self_add(rb_ctx *ctx, double *c, const double *p, int64_t n) { return launch_axpy<3>(ctx, c, p, 0.0, 0.0, n); }
extern "C" int rb_self_sub(rb_ctx *ctx, double *c, const double *p, int64_t n) { return launch_axpy<4>(ctx, c, p, 0.0, 0.0, n); }

// y[i*inc] = beta == 0 ? 0 : beta*y[i*inc]
__global__ void __launch_bounds__(256) rb_scale_or_zero_kernel(double *__restrict__ y, i64 n, i64 inc, double beta)
{
    i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        y[i * inc] = (beta == 0.0) ? 0.0 : beta * y[i * inc];
}
int rb_scale_or_zero(rb_ctx *ctx, double *y, i64 n, i64 inc, double beta)
{
    if (n <= 0 || beta == 1.0) return RB_OK;
    i64 blocks = rb_cdiv(n, 256);
    i64 cap = (i64)ctx->num_sms * 16;
    if (blocks > cap) blocks = cap;
    rb_scale_or_zero_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(y, n, inc, beta);
    RB_LAUNCHED(ctx);
    return RB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Counter-based synthetic inputs; same splitmix64 finaliser as oracle/rest_oracle.c.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double rb_synth(uint64_t seed, uint64_t idx, double scale)
{
    uint64_t z = seed + 0x9E3779B97F4A7C15ULL * (idx + 1ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z ^= z >> 31;
    double u = __dmul_rn((double)(z >> 11), 1.0 / 9007199254740992.0);
    return __dmul_rn(__dsub_rn(__dmul_rn(2.0, u), 1.0), scale);
}

__global__ void __launch_bounds__(256) rb_fill_linear_kernel(double *__restrict__ v, i64 n, uint64_t seed,
                                                             uint64_t idx0, double scale)
{
    i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        v[i] = rb_synth(seed, idx0 + (uint64_t)i, scale);
}

__global__ void __launch_bounds__(256) rb_fill_ri3ao_symm_kernel(double *__restrict__ a, i64 nb, i64 p_lo, i64 nslab,
                                                                 uint64_t seed, double scale)
{
    i64 n2 = nb * nb;
    i64 total = n2 * nslab;
    i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        i64 p = t / n2, r = t - p * n2;
        i64 nu = r / nb, mu = r - nu * nb;
        i64 lo = mu < nu ? mu : nu, hi = mu < nu ? nu : mu;
        a[t] = rb_synth(seed, (uint64_t)(lo + hi * nb + (p + p_lo) * n2), scale);
    }
}

extern "C" int rb_fill_linear(rb_ctx *ctx, double *v, int64_t n, uint64_t seed, uint64_t idx0, double scale)
{
    RB_REQUIRE(ctx && n >= 0, "rb_fill_linear: bad arguments");
    if (n == 0) return RB_OK;
    RB_CUDA(cudaSetDevice(ctx->device));
    i64 blocks = rb_cdiv(n, 256);
    i64 cap = (i64)ctx->num_sms * 16;
    if (blocks > cap) blocks = cap;
    rb_fill_linear_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(v, n, seed, idx0, scale);
    RB_LAUNCHED(ctx);
    return RB_OK;
}

extern "C" int rb_fill_ri3ao_symm(rb_ctx *ctx, double *a, int64_t nb, int64_t p_lo, int64_t p_hi, uint64_t seed,
                                  double scale)
{
    RB_REQUIRE(ctx && nb >= 0 && p_hi >= p_lo, "rb_fill_ri3ao_symm: bad arguments");
    i64 total = nb * nb * (p_hi - p_lo);
    if (total == 0) return RB_OK;
    RB_CUDA(cudaSetDevice(ctx->device));
    i64 blocks = rb_cdiv(total, 256);
    i64 cap = (i64)ctx->num_sms * 16;
    if (blocks > cap) blocks = cap;
    rb_fill_ri3ao_symm_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(a, nb, p_lo, p_hi - p_lo, seed, scale);
    RB_LAUNCHED(ctx);
    return RB_OK;
}
