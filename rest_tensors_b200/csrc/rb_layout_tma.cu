// rb_layout_tma.cu -- TMA (cp.async.bulk.tensor) forms of the bit-exact tile-movement kernels of rb_layout.cu:
//   strided sub-box copies (copy_mm / copy_rr / the unit-stride modes of copy_mr / copy_rm, transpose_ikj)
//   batched 2-D transposes (MatrixFull::transpose, RIFull transpose_jik / jki / kji)
// Reference semantics: restmatr.f90:197-285 (copy_*), src/ri.rs:227-294 (transposes), src/matrix/matrixfull.rs:579-614.
//
// Both are persistent kernels, one CTA per SM, whose global traffic is bulk-tensor transactions only: whole boxes (<= 32 KB)
// are loaded into a shared-memory ring by cp.async.bulk.tensor (completion on mbarriers) and written back by
// cp.async.bulk.tensor stores (bulk groups), so there is no per-thread address arithmetic, no partial-sector access at
// ragged edges (TMA clips boxes at the tensor bounds in both directions) and >= 128 KB per SM stays in flight.
//   copy      : one elected thread per CTA drives the whole pipeline (load box -> wait -> store box from the same buffer).
//   transpose : a producer warp streams 64 x 64 tiles in (four 128B-swizzled boxes of 16 x 64 each), eight consumer warps
//               move them to a second buffer transposed (LDS.128 along r, conflict-free thanks to the swizzle; STS.64
//               along c, contiguous per warp), one elected consumer stores that buffer with a single 64 x 64 box.
// TMA needs 16-byte aligned bases and strides: operands with an odd leading dimension or an 8-byte-aligned base keep the
// plain-load kernels of rb_layout.cu (the callers fall back when rb_tma_* returns RB_TMA_NOT_ELIGIBLE).
// MEASURED (profiles/r02_hbm_kernels.md): these kernels move 5.0-5.7 TB/s, no more than the 16-byte plain-load kernels they
// were meant to replace, while 32-byte LDG/STG kernels (rb_layout.cu) reach 6.2-6.6 TB/s; the bulk-tensor path is therefore
// opt-in (rb_ctx_set_layout_path(ctx, 1) or REST_B200_LAYOUT_TMA=1) and kept for comparison and for its tests.
#include "rb_common.cuh"

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *tm, uint32_t bar, int c0, int c1, int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *tm, uint32_t src, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(tm), "r"(src), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// A logical (i, j, k) box index space; each side's tensor map lists (j, k) in ascending-stride order, `swap` says which.
struct TileSpace {
    i64 tiles_i, tiles_j, total; // total = tiles_i * tiles_j * nk
    int bi, bj;                  // box extent along i and j (k extent is 1)
    int swap_src, swap_dst;      // 1: the map's dimensions are (i, k, j)
};

__device__ __forceinline__ void decode_tile(const TileSpace &p, i64 t, int &ci, int &cj, int &ck)
{
    const i64 ti = t % p.tiles_i, r = t / p.tiles_i;
    const i64 tj = r % p.tiles_j;
    ci = (int)(ti * p.bi); cj = (int)(tj * p.bj); ck = (int)(r / p.tiles_j);
}

// ---- copy: load box -> store box, one thread per CTA ------------------------------------------------------------------
constexpr int CP_STAGES = 6;
constexpr int CP_STAGE_BYTES = 32768;
constexpr int CP_SMEM = CP_STAGES * CP_STAGE_BYTES + 1024 + 64;

__global__ void __launch_bounds__(32, 1)
rb_tma_copy_kernel(const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmD, const TileSpace p)
{
    extern __shared__ uint8_t smem_raw[];
    if (threadIdx.x != 0) return;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + CP_STAGES * CP_STAGE_BYTES;
    for (int s = 0; s < CP_STAGES; ++s) mbar_init(bars + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmS) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmD) : "memory");
    const i64 first = blockIdx.x, step = gridDim.x;
    const i64 count = first < p.total ? (p.total - first + step - 1) / step : 0;
    const uint32_t box_bytes = (uint32_t)p.bi * (uint32_t)p.bj * 8u;
    auto load = [&](i64 n) {
        const int s = (int)(n % CP_STAGES);
        int ci, cj, ck;
        decode_tile(p, first + n * step, ci, cj, ck);
        mbar_expect_tx(bars + 8 * s, box_bytes);
        tma_load_3d(base + s * CP_STAGE_BYTES, &tmS, bars + 8 * s, ci, p.swap_src ? ck : cj, p.swap_src ? cj : ck);
    };
    for (i64 n = 0; n < CP_STAGES - 1 && n < count; ++n) load(n);
    for (i64 n = 0; n < count; ++n) {
        const int s = (int)(n % CP_STAGES);
        mbar_wait(bars + 8 * s, (uint32_t)((n / CP_STAGES) & 1));
        int ci, cj, ck;
        decode_tile(p, first + n * step, ci, cj, ck);
        tma_store_3d(&tmD, base + s * CP_STAGE_BYTES, ci, p.swap_dst ? ck : cj, p.swap_dst ? cj : ck);
        bulk_commit();
        if (n + CP_STAGES - 1 < count) {
            bulk_wait_read<1>(); // store n-1 has read its buffer: that buffer takes load n + CP_STAGES - 1
            load(n + CP_STAGES - 1);
        }
    }
    bulk_wait_all();
}

// ---- transpose: out[c, r, b] = in[r, c, b] over 64 x 64 tiles ---------------------------------------------------------
constexpr int TR = 64;                       // tile edge
constexpr int TR_TILE_BYTES = TR * TR * 8;   // 32 KB
constexpr int TR_LOAD_STAGES = 4, TR_STORE_BUFS = 2;
constexpr int TR_CONSUMER_WARPS = 8;
constexpr int TR_THREADS = (TR_CONSUMER_WARPS + 1) * 32;
constexpr int TR_SMEM = (TR_LOAD_STAGES + TR_STORE_BUFS) * TR_TILE_BYTES + 1024 + 128;

struct TransposeSpace {
    i64 tiles_r, tiles_c, total; // total = tiles_r * tiles_c * nbatch
    int swap_in, swap_out;       // 1: the map lists the batch before the slow matrix dimension
};

__global__ void __launch_bounds__(TR_THREADS, 1)
rb_tma_transpose_kernel(const __grid_constant__ CUtensorMap tmI, const __grid_constant__ CUtensorMap tmO, const TransposeSpace p)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t store_base = base + TR_LOAD_STAGES * TR_TILE_BYTES;
    const uint32_t bars = store_base + TR_STORE_BUFS * TR_TILE_BYTES; // full[L], empty[L]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < TR_LOAD_STAGES; ++s) {
            mbar_init(bars + 8 * s, 1);
            mbar_init(bars + 8 * (TR_LOAD_STAGES + s), TR_CONSUMER_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const i64 first = blockIdx.x, step = gridDim.x;
    const i64 count = first < p.total ? (p.total - first + step - 1) / step : 0;

    if (warp == TR_CONSUMER_WARPS) { // ---- producer warp: one lane streams the tiles in
        if (lane != 0) return;
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmI) : "memory");
        for (i64 n = 0; n < count; ++n) {
            const int s = (int)(n % TR_LOAD_STAGES);
            const uint32_t ph = (uint32_t)((n / TR_LOAD_STAGES) & 1);
            mbar_wait(bars + 8 * (TR_LOAD_STAGES + s), ph ^ 1u);
            const i64 t = first + n * step;
            const i64 tr = t % p.tiles_r, rest = t / p.tiles_r;
            const int r0 = (int)(tr * TR), c0 = (int)((rest % p.tiles_c) * TR), b = (int)(rest / p.tiles_c);
            mbar_expect_tx(bars + 8 * s, TR_TILE_BYTES);
#pragma unroll
            for (int rb = 0; rb < TR / 16; ++rb) // box {16 r, 64 c}: 128-byte rows, SWIZZLE_128B
                tma_load_3d(base + s * TR_TILE_BYTES + rb * (16 * TR * 8), &tmI, bars + 8 * s, r0 + rb * 16,
                            p.swap_in ? b : c0, p.swap_in ? c0 : b);
        }
        return;
    }

    // ---- consumers: warp w moves the row pairs 4w .. 4w+3 (rows 8w .. 8w+7 of the input tile) of all 64 columns
    if (threadIdx.x == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmO) : "memory");
    for (i64 n = 0; n < count; ++n) {
        const int s = (int)(n % TR_LOAD_STAGES), u = (int)(n % TR_STORE_BUFS);
        if (threadIdx.x == 0 && n >= TR_STORE_BUFS) bulk_wait_read<TR_STORE_BUFS - 1>(); // store n - BUFS has read buffer u
        asm volatile("bar.sync 1, %0;" ::"n"(TR_CONSUMER_WARPS * 32) : "memory");
        mbar_wait(bars + 8 * s, (uint32_t)((n / TR_LOAD_STAGES) & 1));
        const uint32_t lbase = base + s * TR_TILE_BYTES, sbase = store_base + u * TR_TILE_BYTES;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k2 = warp * 4 + q;            // row pair: r = 2 k2, 2 k2 + 1
            const int rb = k2 >> 3, ch = k2 & 7;    // 16-row box, 16-byte chunk inside the 128-byte row
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c = lane + 32 * h;
                double2 v;
                asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];"
                             : "=d"(v.x), "=d"(v.y)
                             : "r"(lbase + rb * (16 * TR * 8) + c * 128 + ((ch ^ (c & 7)) << 4)));
                // out tile: [r][c], c contiguous (the 64 x 64 store box)
                asm volatile("st.shared.f64 [%0], %1;" ::"r"(sbase + (2 * k2) * (TR * 8) + c * 8), "d"(v.x) : "memory");
                asm volatile("st.shared.f64 [%0], %1;" ::"r"(sbase + (2 * k2 + 1) * (TR * 8) + c * 8), "d"(v.y) : "memory");
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + 8 * (TR_LOAD_STAGES + s)); // the load stage is free again
        fence_proxy_async();                                        // generic-proxy writes -> visible to the TMA store
        asm volatile("bar.sync 2, %0;" ::"n"(TR_CONSUMER_WARPS * 32) : "memory");
        if (threadIdx.x == 0) {
            const i64 t = first + n * step;
            const i64 tr = t % p.tiles_r, rest = t / p.tiles_r;
            const int r0 = (int)(tr * TR), c0 = (int)((rest % p.tiles_c) * TR), b = (int)(rest / p.tiles_c);
            tma_store_3d(&tmO, sbase, c0, p.swap_out ? b : r0, p.swap_out ? r0 : b);
            bulk_commit();
        }
    }
    if (threadIdx.x == 0) bulk_wait_all();
}

bool stride_ok(i64 s) { return s > 0 && (s & 1) == 0 && s * 8 < (1LL << 40); }

// 3-D FP64 tensor map {d0 (unit stride), d1 @ s1, d2 @ s2}, s1 <= s2 (elements)
int encode3(rb_ctx *ctx, CUtensorMap *tm, const double *basep, i64 d0, i64 d1, i64 s1, i64 d2, i64 s2, int b0, int b1, int b2,
            bool swizzle128)
{
    cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
    cuuint64_t strides[2] = {(cuuint64_t)s1 * 8, (cuuint64_t)s2 * 8};
    cuuint32_t box[3] = {(cuuint32_t)b0, (cuuint32_t)b1, (cuuint32_t)b2};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = ctx->encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void *)basep, dims, strides, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : 1;
}

// one side of a (i, j, k) box space: unit stride along i, strides sj / sk along j / k; returns the swap flag through `swap`
int encode_side(rb_ctx *ctx, CUtensorMap *tm, const double *basep, i64 ni, i64 nj, i64 sj, i64 nk, i64 sk, int bi, int bj,
                bool swizzle128, int *swap)
{
    if (nj == 1 && !stride_ok(sj)) sj = (ni + 1) & ~(i64)1; // the stride of an extent-1 dimension is never used
    if (nk == 1 && !stride_ok(sk)) sk = sj * nj;
    if (nk == 1 && sk < sj) sk = sj * nj;
    if (!stride_ok(sj) || !stride_ok(sk) || sj < ni || sk < ni) return 1;
    if (sj <= sk) { *swap = 0; return encode3(ctx, tm, basep, ni, nj, sj, nk, sk, bi, bj, 1, swizzle128); }
    *swap = 1;
    return encode3(ctx, tm, basep, ni, nk, sk, nj, sj, bi, 1, bj, swizzle128);
}

} // namespace

// dst[i + j*dj + k*dk] = src[i + j*sj + k*sk] (bases already offset).  RB_TMA_NOT_ELIGIBLE: the caller runs the plain kernel.
int rb_tma_copy3d(rb_ctx *ctx, const double *s, i64 sj, i64 sk, double *d, i64 dj, i64 dk, i64 ni, i64 nj, i64 nk)
{
    if (ctx->layout_path != 1 || !ctx->encode_tiled) return RB_TMA_NOT_ELIGIBLE;
    if ((((uintptr_t)s) | ((uintptr_t)d)) & 15) return RB_TMA_NOT_ELIGIBLE;
    if (ni >= (1LL << 31) || nj >= (1LL << 31) || nk >= (1LL << 31)) return RB_TMA_NOT_ELIGIBLE;
    if (ni * nj * nk < (1LL << 16)) return RB_TMA_NOT_ELIGIBLE; // < 512 KB: one wave of the plain kernel is quicker to start
    // Measured on B200: a bulk-tensor STORE clips the unit-stride extent at 16-byte granularity -- with an odd extent of
    // doubles it also writes element [extent] (the zero the load filled in).  Loads clip exactly.
    if (ni & 1) return RB_TMA_NOT_ELIGIBLE;
    // boxes of <= 4096 doubles: the i extent is cut into equal even pieces of <= 256, the j extent fills the box
    const i64 pieces_i = rb_cdiv(ni, 256);
    i64 bi = rb_cdiv(ni, pieces_i);
    bi += bi & 1;
    i64 bj_max = 4096 / bi;
    if (bj_max > 256) bj_max = 256;
    if (bj_max > nj) bj_max = nj;
    i64 bj = rb_cdiv(nj, rb_cdiv(nj, bj_max));
    // enough boxes for every SM's ring: shrink the boxes (not below 4 KB) while there are fewer than 4 per SM
    while (bj > 1 && bi * bj * 8 > 4096 && rb_cdiv(ni, bi) * rb_cdiv(nj, bj) * nk < (i64)4 * ctx->num_sms) bj = (bj + 1) / 2;
    TileSpace p;
    p.bi = (int)bi; p.bj = (int)bj;
    p.tiles_i = rb_cdiv(ni, bi); p.tiles_j = rb_cdiv(nj, bj);
    p.total = p.tiles_i * p.tiles_j * nk;
    CUtensorMap tmS, tmD;
    if (encode_side(ctx, &tmS, s, ni, nj, sj, nk, sk, p.bi, p.bj, false, &p.swap_src)) return RB_TMA_NOT_ELIGIBLE;
    if (encode_side(ctx, &tmD, d, ni, nj, dj, nk, dk, p.bi, p.bj, false, &p.swap_dst)) return RB_TMA_NOT_ELIGIBLE;
    static bool attr_set[64] = {false};
    const int dev = ctx->device & 63;
    if (!attr_set[dev]) {
        RB_CUDA(cudaFuncSetAttribute(rb_tma_copy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CP_SMEM));
        attr_set[dev] = true;
    }
    const int grid = (int)(p.total < ctx->num_sms ? p.total : ctx->num_sms);
    rb_tma_copy_kernel<<<grid, 32, CP_SMEM, ctx->stream>>>(tmS, tmD, p);
    RB_LAUNCHED(ctx);
    ctx->tma_layout_launches++;
    return RB_OK;
}

// out[c + r*ors + b*obs] = in[r + c*ics + b*ibs]
int rb_tma_transpose(rb_ctx *ctx, const double *in, i64 ics, i64 ibs, double *out, i64 ors, i64 obs, i64 nr, i64 nc, i64 nbatch)
{
    if (ctx->layout_path != 1 || !ctx->encode_tiled) return RB_TMA_NOT_ELIGIBLE;
    if ((((uintptr_t)in) | ((uintptr_t)out)) & 15) return RB_TMA_NOT_ELIGIBLE;
    if (nr >= (1LL << 31) || nc >= (1LL << 31) || nbatch >= (1LL << 31)) return RB_TMA_NOT_ELIGIBLE;
    if (nr * nc * nbatch < (1LL << 16)) return RB_TMA_NOT_ELIGIBLE;
    if (nc & 1) return RB_TMA_NOT_ELIGIBLE; // stores clip the unit-stride extent (nc on the output side) in 16-byte units
    TransposeSpace p;
    p.tiles_r = rb_cdiv(nr, TR); p.tiles_c = rb_cdiv(nc, TR);
    p.total = p.tiles_r * p.tiles_c * nbatch;
    CUtensorMap tmI, tmO;
    if (encode_side(ctx, &tmI, in, nr, nc, ics, nbatch, ibs, 16, TR, true, &p.swap_in)) return RB_TMA_NOT_ELIGIBLE;
    if (encode_side(ctx, &tmO, out, nc, nr, ors, nbatch, obs, TR, TR, false, &p.swap_out)) return RB_TMA_NOT_ELIGIBLE;
    static bool attr_set[64] = {false};
    const int dev = ctx->device & 63;
    if (!attr_set[dev]) {
        RB_CUDA(cudaFuncSetAttribute(rb_tma_transpose_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TR_SMEM));
        attr_set[dev] = true;
    }
    const int grid = (int)(p.total < ctx->num_sms ? p.total : ctx->num_sms);
    rb_tma_transpose_kernel<<<grid, TR_THREADS, TR_SMEM, ctx->stream>>>(tmI, tmO, p);
    RB_LAUNCHED(ctx);
    ctx->tma_layout_launches++;
    return RB_OK;
}
