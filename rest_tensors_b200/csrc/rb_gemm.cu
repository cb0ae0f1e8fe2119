// rb_gemm.cu -- FP64 GEMM core for sm_100a.
//
// sm_100a has no tcgen05 FP64 MMA kind; the FP64 tensor path is warp-level mma.sync.m8n8k4.f64, which ptxas
// lowers to DMMA.8x8x4 (measured 37.0 TFLOP/s on this pool's B200 = one DMMA per 16 clk per SM sub-partition).
// The fast kernel is a persistent, warp-specialised pipeline, one CTA of 384 threads per SM:
//
//   producer warp : one elected lane owns all index arithmetic.  It takes work items from a dynamic scheduler (first
//                   item static, then a global atomic counter fetched one tile ahead; the last CTA re-arms the
//                   counter), decodes (batch, tile, k-split), picks the warp grid for ragged tiles, publishes a tile
//                   descriptor in a small shared-memory queue, and issues TMA (cp.async.bulk.tensor.3d; third
//                   coordinate = batch) box loads of the A and B tiles into a 3-stage ring of 64 KB stages, completion
//                   signalled on per-stage "full" mbarriers.  setmaxnreg moves registers from this warpgroup (56) to
//                   the consumers (224).
//   8 consumer warps: each owns up to 2 x 4 blocks of 16 x 16 (32 x 64 of the 128 x 128 CTA tile = 32 DMMA 8x8
//                   accumulator tiles = 128 registers), reads fragments with conflict-free LDS.128, and releases the
//                   stage on a per-stage "empty" mbarrier.  The descriptor arrives with the tile's first stage, so
//                   the consumers do no integer divisions; their per-tile code is ~300 instructions.
//
// The DMMA computes the TRANSPOSED 8x8 tile (A-operand = the B-matrix fragment, B-operand = the A-matrix fragment):
// its accumulator fragment then gives a thread two consecutive rows m = 2t, 2t+1 of column n = g, i.e. 16-byte stores
// along M straight from the accumulator registers (no shuffles, no selects), 64 contiguous bytes per 4 lanes.
//
// Operand layouts in shared memory (chosen so that every fragment read is one LDS.128 with no bank conflicts):
//   K-major operand (op='T' for A, 'N' for B; element (r,k) at k + r*ld):  TMA box {8 k, 128 rows}, no swizzle,
//       smem [k8-block][row][8 doubles]; lanes (g, t) read row g (tile e) / 8+g (tile o), bytes 16t..16t+15: a
//       quarter-warp covers two adjacent 64-byte rows = one 128-byte bank window.  The 16 bytes are k = 2t, 2t+1 ->
//       the two DMMAs of a k8 block use the k-sets {0,2,4,6} and {1,3,5,7} (the summation index may be permuted
//       freely as long as A and B agree).
//   MN-major operand (op='N' for A, 'T' for B; element (r,k) at r + k*ld): TMA box {16 rows, 32 k}, SWIZZLE_128B,
//       smem [16-row block][k][16 doubles ^ swizzle]; a lane reads rows (2g, 2g+1) of a block at k = 2t+j, so the
//       e / o DMMA tiles are the even / odd rows.
//
// K is consumed at k8 granularity: full 32-deep stages run a fully unrolled body, the last stage only its valid k8
// blocks (TMA zero-fills out-of-range rows / k, so edge tiles need no special loads).  Ragged tiles are worked on at
// 16 x 16 block granularity: the 8 warps form a gm x gn grid chosen per tile so that the busiest warp has as few
// blocks as possible, and the k loop is specialised on (blocks_m, blocks_n) per warp OUTSIDE the loop -- every DMMA is
// unpredicated and the hot loop has no dispatch.
//
// Work item = (batch, tile_m, tile_n, k-split); grouped rasterisation (8 tile rows per group) keeps the operands of
// concurrently running CTAs in L2; `tri` enumerates only the tiles of one triangle (SYRK); split-K writes partials
// that a second kernel reduces in a fixed order (deterministic, no FP64 atomics; the dynamic scheduler only changes
// WHICH CTA computes a tile, never how).
//
// Diagonal tiles of a triangular product (SYRK and friends) are worked on as a LIST of 16 x 16 blocks: the 36 blocks on or
// above the diagonal of the 8 x 8 block grid are dealt round-robin to the 8 warps (5 or 4 each, 9 per SM sub-partition),
// so a diagonal tile costs 5/8 of a full tile instead of 8/8 for 36/64 of its work; when both operands are the same
// matrix only one operand tile is loaded.  Batched products whose M is not a multiple of 128 pack their 16-row blocks
// across batches ("packed M", see GemmParams).
//
// A second, generic kernel (plain loads, any alignment / leading dimension) covers operands TMA cannot describe.
#include "rb_common.cuh"
#include <cmath>
#include <vector>

namespace {

constexpr int BM = 128, BN = 128, BK = 32;
constexpr int STAGES = 3;
constexpr int A_TILE_BYTES = BM * BK * 8;
constexpr int B_TILE_BYTES = BN * BK * 8;
constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
constexpr int INFO_SLOTS = 8;  // tile-descriptor queue depth (> STAGES + 1: the producer is at most STAGES stages ahead)
constexpr int INFO_INTS = 8;   // tm, tn, batch, split, full 32-deep stages, tail k8 blocks, warp-grid code, unused
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 64 /*barriers*/ + INFO_SLOTS * INFO_INTS * 4;
constexpr int NUM_CONSUMER_WARPS = 8;
constexpr int NUM_THREADS = (NUM_CONSUMER_WARPS + 4) * 32; // 2 consumer warpgroups + 1 producer warpgroup
constexpr int GROUP_M = 8;

// Stream-K tables (host-built, passed by value): CTA c works on the steps from (cta_tile[c], cta_step[c]) up to, not including,
// (cta_tile[c+1], cta_step[c+1]) of the tile-major (tile, 32-deep k step) space; tile t is covered by tile_pieces[t] consecutive CTAs
// starting with tile_first[t].
constexpr int SK_MAX_TILES = 592, SK_MAX_CTAS = 160;
struct SkTables {
    unsigned cta_step[SK_MAX_CTAS + 1];
    unsigned short cta_tile[SK_MAX_CTAS + 1];
    unsigned char tile_first[SK_MAX_TILES], tile_pieces[SK_MAX_TILES];
};

struct GemmParams {
    i64 m, n, k;
    i64 batch;
    i64 tiles_m, tiles_n;
    i64 tiles_per_batch; // tiles_m*tiles_n or the triangular count
    i64 splits, kper;    // k per split (multiple of BK)
    i64 total_items;
    double alpha, beta;
    double *c;
    i64 ldc, stride_c;
    int tri;             // 0 full, 1 upper (row<=col), 2 lower (row>=col)
    double *partial;     // split-K partials [split][batch][n][ldp]
    i64 ldp;
    int a_batched, b_batched; // 0 => operand shared by all batches (TMA batch coordinate 0)
    unsigned long long *sched; // [0] next work item - gridDim.x, [1] CTAs done (both 0 between launches)
    // packed M (batched, A MN-major, B shared by all batches): the M index runs over 16-row blocks of (batch, row block),
    // 8 consecutive blocks make a tile, so a tile may straddle batches and only the last block of a batch is ragged
    // (m = 600: 38 blocks per batch instead of 5 tiles = 40; m = 264: 17 instead of 24).
    int pack_m;
    i64 bpb, total_blocks;     // 16-row blocks per batch, batch * bpb
    int same_ab;               // tri != 0 and A, B are the same matrix: diagonal tiles load one operand tile only
    // Stream-K: every CTA takes an equal share of the COST of the tile x k-step space (a step of a ragged or diagonal tile weighs what
    // it was measured to cost); a tile that falls into one CTA is written straight to C, the others leave one partial per piece (piece
    // index = CTA - first CTA of the tile) and rb_streamk_reduce_kernel sums them in piece order.  Static partition: no scheduler, every
    // piece is reproducible.  `splits` holds the largest piece count (the partial buffer is [piece][batch][n][ldp] as for split-K).
    int stream_k;
    i64 sk_ksteps;             // 32-deep k steps per tile
    SkTables sk;
};

// ---- PTX wrappers ------------------------------------------------------------------------------------------
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
__device__ __forceinline__ double2 lds128(uint32_t addr)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}

// work item -> (batch, tm, tn, split)
__device__ __forceinline__ void decode_tile(const GemmParams &p, i64 tile, i64 &b, i64 &tm, i64 &tn)
{
    b = tile / p.tiles_per_batch;
    i64 t = tile - b * p.tiles_per_batch;
    if (p.tri == 0) {
        i64 per_group = GROUP_M * p.tiles_n;
        i64 grp = t / per_group;
        i64 first_m = grp * GROUP_M;
        i64 gsz = p.tiles_m - first_m < GROUP_M ? p.tiles_m - first_m : GROUP_M;
        i64 r = t - grp * per_group;
        tm = first_m + r % gsz;
        tn = r / gsz;
    } else {
        // triangular enumeration: t = hi(hi+1)/2 + lo, lo <= hi
        i64 hi = (i64)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
        while (hi * (hi + 1) / 2 > t) --hi;
        while ((hi + 1) * (hi + 2) / 2 <= t) ++hi;
        i64 lo = t - hi * (hi + 1) / 2;
        if (p.tri == 1) { tm = lo; tn = hi; } else { tm = hi; tn = lo; }
    }
}
__device__ __forceinline__ void decode_item(const GemmParams &p, i64 item, i64 &b, i64 &tm, i64 &tn, i64 &sp)
{
    sp = item % p.splits;
    decode_tile(p, item / p.splits, b, tm, tn);
}

// One k8 block (two DMMA k4 steps) for a warp that owns NI x NJ 16x16 blocks (compile-time, so that every DMMA is
// unpredicated: predicated mma.sync costs a WARPSYNC per group).  12 LDS.128 feed 64 DMMAs in the full 2 x 4 case.
template <bool A_K, bool B_K, int NI, int NJ>
__device__ __forceinline__ void consume_kb(double (&acc)[2][2][4][2][2], uint32_t sa, uint32_t sb,
                                           const uint32_t (&a_off)[2], const uint32_t (&b_off)[2], int kb)
{
    double af[NI][2][2]; // [i][tile e/o][mma 0/1]
    double bf[NJ][2][2]; // [j][tile e/o][mma 0/1]
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        if (A_K) {
#pragma unroll
            for (int pa = 0; pa < 2; ++pa) {
                double2 v = lds128(sa + kb * (BM * 64) + i * (16 * 64) + a_off[pa]);
                af[i][pa][0] = v.x; af[i][pa][1] = v.y;
            }
        } else {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                double2 v = lds128(sa + i * (BK * 128) + kb * (8 * 128) + a_off[q]);
                af[i][0][q] = v.x; af[i][1][q] = v.y;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        if (B_K) {
#pragma unroll
            for (int pb = 0; pb < 2; ++pb) {
                double2 v = lds128(sb + kb * (BN * 64) + j * (16 * 64) + b_off[pb]);
                bf[j][pb][0] = v.x; bf[j][pb][1] = v.y;
            }
        } else {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                double2 v = lds128(sb + j * (BK * 128) + kb * (8 * 128) + b_off[q]);
                bf[j][0][q] = v.x; bf[j][1][q] = v.y;
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int i = 0; i < NI; ++i)
#pragma unroll
            for (int pa = 0; pa < 2; ++pa)
#pragma unroll
                for (int j = 0; j < NJ; ++j)
#pragma unroll
                    for (int pb = 0; pb < 2; ++pb) dmma(acc[i][pa][j][pb], bf[j][pb][q], af[i][pa][q]); // D[n][m]: see header
}

// One k8 block for ONE 16 x 16 block of a warp's block list (diagonal tiles): accumulators live in slot S of the same
// register array (block S <-> acc[S >> 2][.][S & 3][.]).  sa_blk / sb_blk already include the block's offset.
template <bool A_K, bool B_K, int S>
__device__ __forceinline__ void consume_list_block(double (&acc)[2][2][4][2][2], uint32_t sa_blk, uint32_t sb_blk,
                                                   const uint32_t (&a_off)[2], const uint32_t (&b_off)[2], int kb)
{
    double af[2][2], bf[2][2]; // [tile e/o][mma 0/1]
    if (A_K) {
#pragma unroll
        for (int pa = 0; pa < 2; ++pa) { double2 v = lds128(sa_blk + kb * (BM * 64) + a_off[pa]); af[pa][0] = v.x; af[pa][1] = v.y; }
    } else {
#pragma unroll
        for (int q = 0; q < 2; ++q) { double2 v = lds128(sa_blk + kb * (8 * 128) + a_off[q]); af[0][q] = v.x; af[1][q] = v.y; }
    }
    if (B_K) {
#pragma unroll
        for (int pb = 0; pb < 2; ++pb) { double2 v = lds128(sb_blk + kb * (BN * 64) + b_off[pb]); bf[pb][0] = v.x; bf[pb][1] = v.y; }
    } else {
#pragma unroll
        for (int q = 0; q < 2; ++q) { double2 v = lds128(sb_blk + kb * (8 * 128) + b_off[q]); bf[0][q] = v.x; bf[1][q] = v.y; }
    }
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int pa = 0; pa < 2; ++pa)
#pragma unroll
            for (int pb = 0; pb < 2; ++pb) dmma(acc[S >> 2][pa][S & 3][pb], bf[pb][q], af[pa][q]);
}

template <bool A_K, bool B_K, int NB>
__device__ __forceinline__ void consume_list_kb(double (&acc)[2][2][4][2][2], uint32_t sa, uint32_t sb, const uint32_t (&la)[5],
                                                const uint32_t (&lb)[5], const uint32_t (&a_off)[2], const uint32_t (&b_off)[2],
                                                int kb)
{
    consume_list_block<A_K, B_K, 0>(acc, sa + la[0], sb + lb[0], a_off, b_off, kb);
    if (NB > 1) consume_list_block<A_K, B_K, 1>(acc, sa + la[1], sb + lb[1], a_off, b_off, kb);
    if (NB > 2) consume_list_block<A_K, B_K, 2>(acc, sa + la[2], sb + lb[2], a_off, b_off, kb);
    if (NB > 3) consume_list_block<A_K, B_K, 3>(acc, sa + la[3], sb + lb[3], a_off, b_off, kb);
    if (NB > 4) consume_list_block<A_K, B_K, 4>(acc, sa + la[4], sb + lb[4], a_off, b_off, kb);
}

// k loop of a diagonal tile for a warp that owns NB (0..5) blocks of the tile's triangle.  b_from_a: both operands are
// the same matrix and only the A tile was loaded; the column-block fragments are read from it.
template <bool A_K, bool B_K, int NB>
__device__ __forceinline__ void run_tile_list(double (&acc)[2][2][4][2][2], uint32_t smem_base, uint32_t bar_base, int &stage,
                                              uint32_t &phase, const uint32_t (&la)[5], const uint32_t (&lb)[5],
                                              const uint32_t (&a_off)[2], const uint32_t (&b_off)[2], int full_steps, int tail_kb,
                                              int lane, bool b_from_a)
{
    for (int ks = 0; ks < full_steps + (tail_kb > 0 ? 1 : 0); ++ks) {
        mbar_wait(bar_base + 8 * stage, phase);
        if (NB > 0) {
            const uint32_t sa = smem_base + stage * STAGE_BYTES, sb = b_from_a ? sa : sa + A_TILE_BYTES;
            if (ks < full_steps) {
#pragma unroll
                for (int kb = 0; kb < BK / 8; ++kb) consume_list_kb<A_K, B_K, (NB > 0 ? NB : 1)>(acc, sa, sb, la, lb, a_off, b_off, kb);
            } else {
#pragma unroll 1
                for (int kb = 0; kb < tail_kb; ++kb) consume_list_kb<A_K, B_K, (NB > 0 ? NB : 1)>(acc, sa, sb, la, lb, a_off, b_off, kb);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_base + 8 * (STAGES + stage));
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
    }
}

// The whole k loop of one tile for a warp with NI x NJ blocks: `full_steps` 32-deep stages (fully unrolled body) and
// one last stage of which only `tail_kb` k8 blocks hold data (K is consumed at k8 granularity, not padded to 32).
// The (NI, NJ) dispatch sits outside this loop, so the hot loop has no per-stage branching.  NI == 0: the warp has
// no valid block in this tile and only keeps the barrier protocol going.
template <bool A_K, bool B_K, int NI, int NJ>
__device__ __forceinline__ void run_tile(double (&acc)[2][2][4][2][2], uint32_t smem_base, uint32_t bar_base, int &stage,
                                         uint32_t &phase, uint32_t a_blk, uint32_t b_blk, const uint32_t (&a_off)[2],
                                         const uint32_t (&b_off)[2], int full_steps, int tail_kb, int lane)
{
    for (int ks = 0; ks < full_steps; ++ks) {
        mbar_wait(bar_base + 8 * stage, phase);
        if (NI > 0) {
            const uint32_t sa = smem_base + stage * STAGE_BYTES + a_blk, sb = smem_base + stage * STAGE_BYTES + A_TILE_BYTES + b_blk;
#pragma unroll
            for (int kb = 0; kb < BK / 8; ++kb) consume_kb<A_K, B_K, (NI > 0 ? NI : 1), NJ>(acc, sa, sb, a_off, b_off, kb);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_base + 8 * (STAGES + stage));
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
    }
    if (tail_kb > 0) {
        mbar_wait(bar_base + 8 * stage, phase);
        if (NI > 0) {
            const uint32_t sa = smem_base + stage * STAGE_BYTES + a_blk, sb = smem_base + stage * STAGE_BYTES + A_TILE_BYTES + b_blk;
#pragma unroll 1
            for (int kb = 0; kb < tail_kb; ++kb) consume_kb<A_K, B_K, (NI > 0 ? NI : 1), NJ>(acc, sa, sb, a_off, b_off, kb);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_base + 8 * (STAGES + stage));
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
    }
}

// Epilogue of one 16-row block of a warp (NJ column blocks), every element inside the matrix: pblk points at
// (first row of the block + row_t, first column of the warp's rectangle + col_t); a thread owns two (A_K) or, across the
// e/o tiles, four (!A_K) consecutive rows of a column -> 16-byte stores, no tests.
template <bool A_K, bool B_K, bool SCALE>
__device__ __forceinline__ void store_lean_row(const double (&acc)[2][2][4][2][2], const int i, double *pblk, i64 ldc,
                                               double alpha, int nj)
{
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (j >= nj) continue;
#pragma unroll
        for (int pb = 0; pb < 2; ++pb) {
            double *pc = pblk + (i64)(j * 16 + (B_K ? 8 * pb : pb)) * ldc;
#pragma unroll
            for (int h = 0; h < 2; ++h) { // A_K: h = pa (rows +8); !A_K: h = cc (rows +2)
                double2 o;
                if (A_K) { o.x = acc[i][h][j][pb][0]; o.y = acc[i][h][j][pb][1]; }
                else { o.x = acc[i][0][j][pb][h]; o.y = acc[i][1][j][pb][h]; }
                if (SCALE) { o.x *= alpha; o.y *= alpha; }
                *reinterpret_cast<double2 *>(pc + (A_K ? 8 * h : 2 * h)) = o;
            }
        }
    }
}

// One 16 x 16 block with every test: matrix bounds (rows < m_lim, cols < n_lim), triangle (tri 1: row <= col, 2: row >=
// col; rows / cols are indices in the matrix cb points at), beta.  (i, j) select the accumulators.
template <bool A_K, bool B_K>
__device__ __forceinline__ void store_block_checked(const double (&acc)[2][2][4][2][2], const int i, const int j, double *cb,
                                                    i64 ldc, i64 row_blk, i64 col_blk, i64 m_lim, i64 n_lim, int tri,
                                                    bool vec_ok, double alpha, double beta, int row_t, int col_t)
{
#pragma unroll
    for (int pb = 0; pb < 2; ++pb) {
        const i64 col = col_blk + col_t + (B_K ? 8 * pb : pb);
        if (col >= n_lim) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h) { // A_K: h = pa (rows +8); !A_K: h = cc (rows +2)
            const double v0 = A_K ? acc[i][h][j][pb][0] : acc[i][0][j][pb][h];
            const double v1 = A_K ? acc[i][h][j][pb][1] : acc[i][1][j][pb][h];
            const i64 row0 = row_blk + row_t + (A_K ? 8 * h : 2 * h);
            double *cp = cb + row0 + col * ldc;
            bool ok0 = row0 < m_lim, ok1 = row0 + 1 < m_lim;
            if (tri == 1) { ok0 = ok0 && row0 <= col; ok1 = ok1 && row0 + 1 <= col; }
            else if (tri == 2) { ok0 = ok0 && row0 >= col; ok1 = ok1 && row0 + 1 >= col; }
            if (ok0 && ok1 && vec_ok) {
                double2 o;
                if (beta == 0.0) { o.x = alpha * v0; o.y = alpha * v1; }
                else {
                    const double2 old = *reinterpret_cast<const double2 *>(cp);
                    o.x = alpha * v0 + beta * old.x; o.y = alpha * v1 + beta * old.y;
                }
                *reinterpret_cast<double2 *>(cp) = o;
            } else {
                if (ok0) cp[0] = (beta == 0.0) ? alpha * v0 : alpha * v0 + beta * cp[0];
                if (ok1) cp[1] = (beta == 0.0) ? alpha * v1 : alpha * v1 + beta * cp[1];
            }
        }
    }
}

// ---- the TMA + DMMA kernel --------------------------------------------------------------------------------------
template <bool A_K, bool B_K, bool PACK>
__global__ void __launch_bounds__(NUM_THREADS, 1)
rb_gemm_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES; // full[STAGES], empty[STAGES]
    const uint32_t info_base = bar_base + 64;                   // tile-descriptor queue
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar_base + 8 * s, 1);
            mbar_init(bar_base + 8 * (STAGES + s), NUM_CONSUMER_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp >= NUM_CONSUMER_WARPS) {
        // ===================== producer warpgroup: hand its registers to the consumers =====================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        if (warp == NUM_CONSUMER_WARPS && lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
            int stage = 0;
            uint32_t phase = 0;
            int tcount = 0;
            // Dynamic tile scheduler: the first item is static, the rest come from a global counter (tiles differ in
            // cost -- ragged edges, triangles -- so a static round-robin leaves SMs idle at the end).  The next item is
            // fetched one tile ahead so that the atomic's latency never sits between two tiles.
            i64 item = blockIdx.x;
            unsigned sk_t = 0, sk_s = 0, sk_t_end = 0, sk_s_end = 0; // stream-K: this CTA's range of (tile, step)
            if (p.stream_k) {
                sk_t = p.sk.cta_tile[blockIdx.x]; sk_s = p.sk.cta_step[blockIdx.x];
                sk_t_end = p.sk.cta_tile[blockIdx.x + 1]; sk_s_end = p.sk.cta_step[blockIdx.x + 1];
            }
            while (p.stream_k ? (sk_t < sk_t_end || (sk_t == sk_t_end && sk_s < sk_s_end)) : item < p.total_items) {
                i64 next = 0;
                i64 b, tm, tn, sp, k_begin, k_end;
                bool whole = false;
                if (p.stream_k) {
                    const i64 s1 = (sk_t == sk_t_end) ? (i64)sk_s_end : p.sk_ksteps;
                    decode_tile(p, (i64)sk_t, b, tm, tn);
                    k_begin = (i64)sk_s * BK;
                    k_end = (s1 * BK < p.k) ? s1 * BK : p.k;
                    sp = (i64)blockIdx.x - (i64)p.sk.tile_first[sk_t]; // piece index inside the tile
                    whole = p.sk.tile_pieces[sk_t] == 1;
                    ++sk_t; sk_s = 0;
                } else {
                    next = (i64)gridDim.x + (i64)atomicAdd(p.sched, 1ULL);
                    decode_item(p, item, b, tm, tn, sp);
                    k_begin = sp * p.kper;
                    k_end = (k_begin + p.kper < p.k) ? k_begin + p.kper : p.k;
                }
                const int ba = p.a_batched ? (int)b : 0, bb = p.b_batched ? (int)b : 0;
                // Tile descriptor for the consumers (they do no index arithmetic of their own).  Valid 16x16 blocks of
                // the tile (edge tiles are ragged; TMA zero-fills the rest) and the gm x gn warp grid that gives the
                // busiest warp the fewest blocks with at most 2 x 4 per warp: 4 x 2 for full tiles, 8 x 1 / 1 x 8 /
                // 2 x 4 for thin ones.  The cost of a tile is thus proportional to its valid blocks, not to 128 x 128.
                const i64 nrem = p.n - tn * BN;
                int mb;
                if (PACK) { const i64 left = p.total_blocks - tm * (BM / 16); mb = left >= BM / 16 ? BM / 16 : (int)left; }
                else { const i64 mrem = p.m - tm * BM; mb = mrem >= BM ? BM / 16 : (int)((mrem + 15) >> 4); }
                const int nbk = nrem >= BN ? BN / 16 : (int)((nrem + 15) >> 4);
                int lgm = 2, rpg = 2, cpg = 4, best = 1 << 30;
#pragma unroll
                for (int lg = 3; lg >= 0; --lg) { // gm = 8, 4, 2, 1; gn = 8 / gm
                    const int r = (mb + (1 << lg) - 1) >> lg, c = (nbk + (8 >> lg) - 1) >> (3 - lg);
                    if (r <= 2 && c <= 4 && r * c < best) { best = r * c; lgm = lg; rpg = r; cpg = c; }
                }
                // diagonal tile of a triangular product: block-list mode (see header); same operand on both sides -> the
                // B tile is not loaded at all
                const bool diag = p.tri != 0 && tm == tn;
                const bool skip_b = diag && p.same_ab;
                const uint32_t grid_code = (uint32_t)lgm | ((uint32_t)rpg << 4) | ((uint32_t)cpg << 8) | ((uint32_t)mb << 12) |
                                           ((uint32_t)nbk << 16) | (diag ? (1u << 20) : 0u) | (whole ? (1u << 21) : 0u);
                const uint32_t tx_bytes = (PACK ? (uint32_t)mb * (16 * BK * 8) : (uint32_t)A_TILE_BYTES) + (skip_b ? 0u : (uint32_t)B_TILE_BYTES);
                // packed M: (batch, first row) of the tile's first 16-row block
                int pk_b0 = 0, pk_r0 = 0;
                if (PACK) { const i64 gb0 = tm * (BM / 16); pk_b0 = (int)(gb0 / p.bpb); pk_r0 = (int)(gb0 - (i64)pk_b0 * p.bpb); }
                bool first = true;
                for (i64 k0 = k_begin; k0 < k_end; k0 += BK) {
                    const uint32_t full = bar_base + 8 * stage, empty = bar_base + 8 * (STAGES + stage);
                    mbar_wait(empty, phase ^ 1u);
                    if (first) { // the arrive below releases the descriptor together with the tile's first stage
                        const uint32_t slot = info_base + (tcount & (INFO_SLOTS - 1)) * INFO_INTS * 4;
                        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(slot), "r"((uint32_t)tm), "r"((uint32_t)tn),
                                     "r"((uint32_t)b), "r"((uint32_t)sp) : "memory");
                        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(slot + 16), "r"((uint32_t)((k_end - k_begin) / BK)),
                                     "r"((uint32_t)(((k_end - k_begin) % BK + 7) >> 3)), "r"(grid_code), "r"(0u) : "memory");
                        first = false;
                    }
                    mbar_expect_tx(full, tx_bytes);
                    const uint32_t sa = smem_base + stage * STAGE_BYTES, sb = sa + A_TILE_BYTES;
                    if (A_K) {
#pragma unroll
                        for (int kb = 0; kb < BK / 8; ++kb)
                            tma_load_3d(sa + kb * (BM * 64), &tmA, full, (int)(k0 + kb * 8), (int)(tm * BM), ba);
                    } else if (PACK) {
                        int bb_ = pk_b0, rr_ = pk_r0;
#pragma unroll
                        for (int blk = 0; blk < BM / 16; ++blk) {
                            if (blk < mb) tma_load_3d(sa + blk * (BK * 128), &tmA, full, rr_ * 16, (int)k0, bb_);
                            if (++rr_ == (int)p.bpb) { rr_ = 0; ++bb_; }
                        }
                    } else {
#pragma unroll
                        for (int blk = 0; blk < BM / 16; ++blk)
                            tma_load_3d(sa + blk * (BK * 128), &tmA, full, (int)(tm * BM + blk * 16), (int)k0, ba);
                    }
                    if (!skip_b) {
                        if (B_K) {
#pragma unroll
                            for (int kb = 0; kb < BK / 8; ++kb)
                                tma_load_3d(sb + kb * (BN * 64), &tmB, full, (int)(k0 + kb * 8), (int)(tn * BN), bb);
                        } else {
#pragma unroll
                            for (int blk = 0; blk < BN / 16; ++blk)
                                tma_load_3d(sb + blk * (BK * 128), &tmB, full, (int)(tn * BN + blk * 16), (int)k0, bb);
                        }
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
                item = next;
                ++tcount;
            }
            { // end-of-work descriptor, released through the next stage's full barrier (no data)
                const uint32_t full = bar_base + 8 * stage, empty = bar_base + 8 * (STAGES + stage);
                mbar_wait(empty, phase ^ 1u);
                const uint32_t slot = info_base + (tcount & (INFO_SLOTS - 1)) * INFO_INTS * 4;
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(slot + 16), "r"(0xffffffffu), "r"(0u), "r"(0u), "r"(0u) : "memory");
                mbar_arrive(full);
            }
            // the last CTA to run dry re-arms the scheduler for the next launch on this stream (stream-K never touched it)
            if (!p.stream_k && atomicAdd(p.sched + 1, 1ULL) == (unsigned long long)gridDim.x - 1ULL) {
                p.sched[0] = 0ULL;
                p.sched[1] = 0ULL;
                __threadfence();
            }
        }
        return;
    }

    // ===================== consumers =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int g = lane >> 2, t = lane & 3;

    // per-thread byte offsets of the fragment reads inside a 16-row block (k8-block / block offsets are added per use)
    uint32_t a_off[2], b_off[2];
    if (A_K) {
        a_off[0] = (uint32_t)(g * 64 + t * 16);        // tile e : row g
        a_off[1] = (uint32_t)((8 + g) * 64 + t * 16);  // tile o : row 8+g
    } else {
        a_off[0] = (uint32_t)((2 * t) * 128 + ((g ^ (2 * t)) << 4));          // k = 2t   : rows (2g, 2g+1)
        a_off[1] = (uint32_t)((2 * t + 1) * 128 + ((g ^ (2 * t + 1)) << 4));  // k = 2t+1
    }
    if (B_K) {
        b_off[0] = (uint32_t)(g * 64 + t * 16);        // tile e : col g
        b_off[1] = (uint32_t)((8 + g) * 64 + t * 16);  // tile o : col 8+g
    } else {
        b_off[0] = (uint32_t)((2 * t) * 128 + ((g ^ (2 * t)) << 4));
        b_off[1] = (uint32_t)((2 * t + 1) * 128 + ((g ^ (2 * t + 1)) << 4));
    }
    // Position of accumulator acc[i][pa][j][pb][cc] inside the warp's rectangle (the DMMA computes the TRANSPOSED 8x8
    // tile, so a thread holds two consecutive rows m = 2t, 2t+1 of column n = g):
    //   row = 16 i + (A_K ? 8 pa + 2t + cc : 4t + 2 cc + pa),   col = 16 j + (B_K ? 8 pb + g : 2g + pb)
    const int row_t = A_K ? 2 * t : 4 * t;
    const int col_t = B_K ? g : 2 * g;

    int stage = 0;
    uint32_t phase = 0;
    int tcount = 0;
    for (;; ++tcount) {
        // the first stage of the tile carries the tile descriptor published by the producer
        mbar_wait(bar_base + 8 * stage, phase);
        uint32_t utm, utn, ub, usp, ufull, utail, ucode, upad;
        {
            const uint32_t slot = info_base + (tcount & (INFO_SLOTS - 1)) * INFO_INTS * 4;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(utm), "=r"(utn), "=r"(ub), "=r"(usp) : "r"(slot) : "memory");
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(ufull), "=r"(utail), "=r"(ucode), "=r"(upad) : "r"(slot + 16) : "memory");
        }
        if (ufull == 0xffffffffu) break; // no more work for this CTA
        const i64 tm = utm, tn = utn, bidx = ub, sp = usp;
        const int full_steps = (int)ufull, tail_kb = (int)utail;
        const int lgm = ucode & 15, rpg = (ucode >> 4) & 15, cpg = (ucode >> 8) & 15, mb = (ucode >> 12) & 15, nbk = (ucode >> 16) & 15;
        double acc[2][2][4][2][2]; // [row block i][row tile e/o][col block j][col tile e/o][2]
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int pa = 0; pa < 2; ++pa)
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int pb = 0; pb < 2; ++pb) { acc[i][pa][j][pb][0] = 0.0; acc[i][pa][j][pb][1] = 0.0; }

        // output matrix of this item
        double *cb;
        i64 ldc, cstride;
        double alpha, beta;
        const bool to_partial = p.splits > 1 && ((ucode >> 21) & 1u) == 0u; // stream-K: a tile owned by one CTA goes straight to C
        if (to_partial) {
            cstride = p.n * p.ldp;
            cb = p.partial + (sp * p.batch + bidx) * cstride;
            ldc = p.ldp; alpha = 1.0; beta = 0.0;
        } else {
            cstride = p.stride_c;
            cb = p.c + bidx * cstride;
            ldc = p.ldc; alpha = p.alpha; beta = p.beta;
        }
        const bool vec_ok = ((ldc & 1) == 0) && ((((uintptr_t)cb) & 15) == 0) && (!PACK || (cstride & 1) == 0);
        const int tri = to_partial ? 0 : p.tri;

        if ((ucode >> 20) & 1u) {
            // ---- diagonal tile of a triangular product: this warp's share of the mb(mb+1)/2 blocks on or above (tri 1)
            //      / below (tri 2) the diagonal, dealt round-robin: block e = warp + 8 s, e = hi(hi+1)/2 + lo, lo <= hi ----
            const int nblocks = mb * (mb + 1) / 2;
            const int nown = warp < nblocks ? (nblocks - warp + 7) >> 3 : 0;
            uint32_t la[5], lb[5];
            int lrow[5], lcol[5];
#pragma unroll
            for (int s = 0; s < 5; ++s) {
                int e = warp + 8 * s, hi = 0;
                if (e >= nblocks) e = 0;
                while ((hi + 1) * (hi + 2) / 2 <= e) ++hi;
                const int lo = e - hi * (hi + 1) / 2;
                lrow[s] = (p.tri == 1) ? lo : hi;
                lcol[s] = (p.tri == 1) ? hi : lo;
                la[s] = (uint32_t)lrow[s] * (A_K ? (16 * 64) : (BK * 128));
                lb[s] = (uint32_t)lcol[s] * (B_K ? (16 * 64) : (BK * 128));
            }
            const bool b_from_a = p.same_ab != 0;
#define RB_RUN_LIST(NB_) run_tile_list<A_K, B_K, NB_>(acc, smem_base, bar_base, stage, phase, la, lb, a_off, b_off, full_steps, tail_kb, lane, b_from_a)
            switch (nown) { // warp-uniform
            case 5: RB_RUN_LIST(5); break;
            case 4: RB_RUN_LIST(4); break;
            case 3: RB_RUN_LIST(3); break;
            case 2: RB_RUN_LIST(2); break;
            case 1: RB_RUN_LIST(1); break;
            default: RB_RUN_LIST(0); break;
            }
#undef RB_RUN_LIST
#pragma unroll
            for (int s = 0; s < 5; ++s) {
                if (s >= nown) continue;
                store_block_checked<A_K, B_K>(acc, s >> 2, s & 3, cb, ldc, tm * BM + lrow[s] * 16, tn * BN + lcol[s] * 16, p.m, p.n,
                                              lrow[s] == lcol[s] ? tri : 0, vec_ok, alpha, beta, row_t, col_t);
            }
            continue;
        }

        const int gi = warp & ((1 << lgm) - 1), gj = warp >> lgm;
        const int rb0 = gi * rpg, cb0 = gj * cpg;
        int ni = mb - rb0; ni = ni < 0 ? 0 : (ni > rpg ? rpg : ni);
        int nj = nbk - cb0; nj = nj < 0 ? 0 : (nj > cpg ? cpg : nj);
        if (ni == 0 || nj == 0) { ni = 0; nj = 0; }
        const uint32_t a_blk = (uint32_t)rb0 * (A_K ? (16 * 64) : (BK * 128));
        const uint32_t b_blk = (uint32_t)cb0 * (B_K ? (16 * 64) : (BK * 128));

#define RB_RUN(NI_, NJ_) run_tile<A_K, B_K, NI_, NJ_>(acc, smem_base, bar_base, stage, phase, a_blk, b_blk, a_off, b_off, full_steps, tail_kb, lane)
        switch (ni * 8 + nj) { // warp-uniform
        case 2 * 8 + 4: RB_RUN(2, 4); break;
        case 2 * 8 + 3: RB_RUN(2, 3); break;
        case 2 * 8 + 2: RB_RUN(2, 2); break;
        case 2 * 8 + 1: RB_RUN(2, 1); break;
        case 1 * 8 + 4: RB_RUN(1, 4); break;
        case 1 * 8 + 3: RB_RUN(1, 3); break;
        case 1 * 8 + 2: RB_RUN(1, 2); break;
        case 1 * 8 + 1: RB_RUN(1, 1); break;
        default: RB_RUN(0, 1); break; // this warp has no valid block in this tile
        }
#undef RB_RUN

        // ---- epilogue, one 16-row block at a time: a thread owns rows (2t, 2t+1) (+8) or (4t .. 4t+3) of each block ->
        //      16-byte stores along M.  A block whose 16 rows and all of the warp's columns lie inside the matrix (the
        //      common case) takes the lean path: no bounds / triangle tests, one IMAD per store, no scaling when alpha == 1.
        //      The epilogue is pure issue overhead for the DMMA pipe, so it is kept as short as possible. ----
        const i64 col_w = tn * BN + cb0 * 16;
        const bool cols_in = col_w + nj * 16 <= p.n;
        const bool lean_ok = vec_ok && tri == 0 && beta == 0.0 && cols_in;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            if (i >= ni) continue;
            double *cm = cb;       // matrix that holds this row block
            i64 row_blk;           // first row of the block inside it
            if (PACK) {
                const i64 gb = tm * (BM / 16) + rb0 + i, bq = gb / p.bpb;
                cm = cb + bq * cstride;
                row_blk = (gb - bq * p.bpb) * 16;
            } else row_blk = tm * BM + (rb0 + i) * 16;
            if (lean_ok && row_blk + 16 <= p.m) {
                double *pblk = cm + (row_blk + row_t) + (col_w + col_t) * ldc;
                if (alpha == 1.0) store_lean_row<A_K, B_K, false>(acc, i, pblk, ldc, alpha, nj);
                else store_lean_row<A_K, B_K, true>(acc, i, pblk, ldc, alpha, nj);
                continue;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (j >= nj) continue;
                store_block_checked<A_K, B_K>(acc, i, j, cm, ldc, row_blk, col_w + j * 16, p.m, p.n, tri, vec_ok, alpha, beta, row_t, col_t);
            }
        }
    }
}

// ---- reduction of the split-K / stream-K partials (fixed order => deterministic) -------------------------------------------
// partial[piece][batch][n][ldp] -> C = alpha * sum_pieces + beta * C.  A thread owns two consecutive rows of a column (16-byte loads of
// the partials, ldp is even; 8 pieces in flight, added in piece order).  SK: the piece count is per tile (SkTables) and tiles owned by
// one CTA were written by the GEMM kernel itself and are skipped.  No integer divisions: (row pair, column, batch) come from the grid.
template <bool SK>
__global__ void __launch_bounds__(256) rb_partial_reduce_kernel(const double *__restrict__ partial, i64 ldp, int splits, i64 m, i64 n,
                                                                i64 batch, double alpha, double beta, double *__restrict__ c, i64 ldc,
                                                                i64 stride_c, int tri, int tiles_m, int tiles_n, int tiles_per_batch,
                                                                const __grid_constant__ SkTables sk)
{
    const i64 split_stride = batch * n * ldp;
    for (i64 b = blockIdx.z; b < batch; b += gridDim.z)
        for (i64 j = blockIdx.y; j < n; j += gridDim.y) {
            const int tn = (int)(j >> 7);
            for (i64 i = 2 * ((i64)blockIdx.x * blockDim.x + threadIdx.x); i < m; i += 2 * (i64)gridDim.x * blockDim.x) {
                bool v0 = true, v1 = i + 1 < m;
                if (tri == 1) { v0 = i <= j; v1 = v1 && i + 1 <= j; }
                else if (tri == 2) { v0 = i >= j; v1 = v1 && i + 1 >= j; }
                if (!v0 && !v1) continue;
                int np = splits;
                if (SK) {
                    const int tm = (int)(i >> 7); // rows i, i + 1 share a tile (i is even)
                    int t;
                    if (tri == 0) {
                        const int grp = tm / GROUP_M, first_m = grp * GROUP_M;
                        const int gsz = tiles_m - first_m < GROUP_M ? tiles_m - first_m : GROUP_M;
                        t = grp * GROUP_M * tiles_n + tn * gsz + (tm - first_m);
                    } else if (tri == 1) t = tn * (tn + 1) / 2 + tm;
                    else t = tm * (tm + 1) / 2 + tn;
                    np = sk.tile_pieces[(int)b * tiles_per_batch + t];
                    if (np == 1) continue;
                }
                const double *pp = partial + (b * n + j) * ldp + i;
                double s0 = 0.0, s1 = 0.0;
                if (i + 1 < m) {
#pragma unroll 8
                    for (int q = 0; q < np; ++q) {
                        const double2 v = *reinterpret_cast<const double2 *>(pp + (i64)q * split_stride);
                        s0 += v.x; s1 += v.y;
                    }
                } else { // last row of an odd m: the pad row of the partials was never written, do not touch it
#pragma unroll 8
                    for (int q = 0; q < np; ++q) s0 += pp[(i64)q * split_stride];
                }
                double *cp = c + b * stride_c + i + j * ldc;
                if (v0) cp[0] = (beta == 0.0) ? alpha * s0 : alpha * s0 + beta * cp[0];
                if (v1) cp[1] = (beta == 0.0) ? alpha * s1 : alpha * s1 + beta * cp[1];
            }
        }
}

// ---- generic kernel: any alignment / leading dimension, plain loads -------------------------------------------------
// 64x64 CTA tile, 4 warps (32x32 each = 4x4 DMMA tiles), BK = 16, smem [k][64+4] (conflict-free LDS.64 fragments).
constexpr int GB = 64, GK = 16, GPAD = 4;

__global__ void __launch_bounds__(128)
rb_gemm_generic_kernel(bool ta, bool tb, i64 m, i64 n, i64 k, double alpha, const double *__restrict__ a, i64 lda,
                       i64 stride_a, const double *__restrict__ b, i64 ldb, i64 stride_b, double beta,
                       double *__restrict__ c, i64 ldc, i64 stride_c, i64 tiles_m, i64 tiles_n, int tri)
{
    __shared__ double As[GK][GB + GPAD];
    __shared__ double Bs[GK][GB + GPAD];
    const i64 tile = blockIdx.x;
    const i64 bidx = blockIdx.y;
    const i64 tm = tile % tiles_m, tn = tile / tiles_m;
    if (tri == 1 && tm > tn) return;
    if (tri == 2 && tm < tn) return;
    const double *ab = a + bidx * stride_a;
    const double *bb = b + bidx * stride_b;
    double *cb = c + bidx * stride_c;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wm = warp & 1, wn = warp >> 1;
    const int g = lane >> 2, t = lane & 3;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

    for (i64 k0 = 0; k0 < k; k0 += GK) {
        // cooperative loads: 64*16 elements per operand, 128 threads -> 8 each
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            int e = threadIdx.x + r * 128;
            int mm, kk;
            if (ta) { kk = e % GK; mm = e / GK; }  // K-major source: consecutive threads walk k
            else { mm = e % GB; kk = e / GB; }
            i64 gm = tm * GB + mm, gk = k0 + kk;
            double v = 0.0;
            if (gm < m && gk < k) v = ta ? ab[gk + gm * lda] : ab[gm + gk * lda];
            As[kk][mm] = v;
            int nn;
            if (!tb) { kk = e % GK; nn = e / GK; } // B 'N' is K-major
            else { nn = e % GB; kk = e / GB; }
            i64 gn = tn * GB + nn;
            gk = k0 + kk;
            v = 0.0;
            if (gn < n && gk < k) v = tb ? bb[gn + gk * ldb] : bb[gk + gn * ldb];
            Bs[kk][nn] = v;
        }
        __syncthreads();
#pragma unroll
        for (int ks = 0; ks < GK; ks += 4) {
            double af[4], bf[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) af[i] = As[ks + t][wm * 32 + i * 8 + g];
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = Bs[ks + t][wn * 32 + j * 8 + g];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma(acc[i][j], af[i], bf[j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                i64 row = tm * GB + wm * 32 + i * 8 + g;
                i64 col = tn * GB + wn * 32 + j * 8 + 2 * t + cc;
                if (row >= m || col >= n) continue;
                if (tri == 1 && row > col) continue;
                if (tri == 2 && row < col) continue;
                double *cp = cb + row + col * ldc;
                *cp = (beta == 0.0) ? alpha * acc[i][j][cc] : alpha * acc[i][j][cc] + beta * (*cp);
            }
}

// C = beta*C over an m x n block (k == 0 or alpha == 0 degenerate cases)
__global__ void __launch_bounds__(256) rb_scale_c_kernel(double *__restrict__ c, i64 m, i64 n, i64 ldc, i64 stride_c,
                                                         i64 batch, double beta, int tri)
{
    i64 total = m * n * batch;
    i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        i64 i = e % m, r = e / m;
        i64 j = r % n, b = r / n;
        if (tri == 1 && i > j) continue;
        if (tri == 2 && i < j) continue;
        double *cp = c + b * stride_c + i + j * ldc;
        *cp = (beta == 0.0) ? 0.0 : beta * (*cp);
    }
}

int encode_map(rb_ctx *ctx, CUtensorMap *tm, const double *base, bool k_major, i64 rows, i64 k, i64 ld, i64 stride,
               i64 batch)
{
    // k_major: dims {k, rows, batch}, box {8, 128, 1}, no swizzle; else dims {rows, k, batch}, box {16, BK, 1}, 128B swizzle
    cuuint64_t dims[3];
    cuuint64_t strides[2];
    cuuint32_t box[3];
    cuuint32_t estr[3] = {1, 1, 1};
    i64 d0 = k_major ? k : rows, d1 = k_major ? rows : k;
    dims[0] = (cuuint64_t)d0; dims[1] = (cuuint64_t)d1; dims[2] = (cuuint64_t)(batch > 0 ? batch : 1);
    strides[0] = (cuuint64_t)ld * 8;
    i64 bs = (batch > 1) ? stride : ld * d1;
    if (bs <= 0) bs = ld * d1;
    strides[1] = (cuuint64_t)bs * 8;
    if (k_major) { box[0] = 8; box[1] = 128; box[2] = 1; }
    else { box[0] = 16; box[1] = BK; box[2] = 1; }
    CUresult r = ctx->encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void *)base, dims, strides, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   k_major ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        rb_set_error("cuTensorMapEncodeTiled failed (%d): dims {%lld,%lld,%lld} ld %lld stride %lld", (int)r,
                     (long long)d0, (long long)d1, (long long)batch, (long long)ld, (long long)bs);
        return RB_ERR_CUDA;
    }
    return RB_OK;
}

bool tma_eligible(const rb_ctx *ctx, const double *p, i64 ld, i64 stride, i64 batch, i64 d0, i64 d1)
{
    if (!ctx->encode_tiled) return false;
    if (((uintptr_t)p) & 15) return false;
    if (ld & 1) return false;
    if (batch > 1 && (stride & 1)) return false;
    if (batch > 1 && stride < 0) return false;
    if (ld * 8 >= (1LL << 40)) return false;
    if (batch > 1 && stride * 8 >= (1LL << 40)) return false;
    if (d0 >= (1LL << 31) || d1 >= (1LL << 31) || batch >= (1LL << 31)) return false; // TMA coordinates are int32
    if (ld < d0) return false;
    return true;
}

// Split-K plan.  When the tiles do not fill whole waves of the chip and K is deep, the split count minimises a makespan
// estimate in k steps: rounds(s) * item cost, where an item costs k/s steps + ~2 fixed (descriptor, epilogue, partial
// store), rounds = ceil(items / SMs) is the number of items the busiest SM takes, and ragged edge tiles weigh what they
// were measured to cost (the kernel works at 16 x 16 block granularity; the dynamic scheduler back-fills cheap items)
// -- e.g. 105 tiles x 1013 steps: s = 7 -> 5 rounds of 147 steps instead of one of 1015; a 264^2 SYRK (3 full + 3 thin
// tiles): s = 48, not the s = 24 that fills one round with half-empty SMs.  Partials are reduced in a fixed order:
// results do not depend on the schedule.  Pure host arithmetic (exported as rb_gemm_plan_splits for the CPU tests).
// packed: the batched product runs with packed M (16-row blocks of all batches enumerated together).
bool pack_m_applies(bool a_k, bool a_batched, bool b_batched, i64 m, i64 batch, int tri)
{
    // A MN-major (one TMA box per 16-row block anyway), B shared by all batches, and something to gain: ragged M
    return !a_k && a_batched && !b_batched && batch > 1 && tri == 0 && (m % BM) != 0 && rb_cdiv(m, 16) < (1LL << 30);
}

i64 plan_splits(i64 m, i64 n, i64 k, i64 batch, int tri, int num_sms, bool packed = false)
{
    const i64 tiles_n = rb_cdiv(n, BN);
    const i64 tiles_m = packed ? rb_cdiv(rb_cdiv(m, 16) * batch, BM / 16) : rb_cdiv(m, BM);
    if (packed) batch = 1;
    const i64 tiles_per_batch = tri ? tiles_m * (tiles_m + 1) / 2 : tiles_m * tiles_n;
    const i64 tiles = tiles_per_batch * batch;
    const i64 ksteps = rb_cdiv(k, BK);
    i64 splits = 1;
    if (tiles <= 0 || num_sms <= 0 || !(tiles < 4 * (i64)num_sms && ksteps >= 8)) return 1;
    const i64 me = packed ? BM : m - (tiles_m - 1) * BM, ne = n - (tiles_n - 1) * BN; // extents of the last tile row / column
    const int mbe = (int)((me + 15) >> 4), nbe = (int)((ne + 15) >> 4);
    const double wm_edge = mbe >= 5 ? 1.0 : mbe >= 3 ? 0.5 : mbe == 2 ? 0.25 : 0.125; // <= 2 block rows per warp row
    const double wn_edge = nbe / 8.0;
    double eff = 0.0, max_w = 0.0;
    for (i64 tn = 0; tn < tiles_n; ++tn)
        for (i64 tm = 0; tm < tiles_m; ++tm) {
            if ((tri == 1 && tm > tn) || (tri == 2 && tm < tn)) continue;
            double w = (tm == tiles_m - 1 ? wm_edge : 1.0) * (tn == tiles_n - 1 ? wn_edge : 1.0);
            if (tri && tm == tn) { // block-list mode: the busiest warp owns ceil(T / 8) of the T = mb(mb+1)/2 blocks
                const int mb = tm == tiles_m - 1 ? mbe : 8;
                w = (double)((mb * (mb + 1) / 2 + 7) / 8) / 8.0;
            }
            if (w < 0.22) w = 0.22; // measured floor of a 1-block-wide tile
            eff += w;
            if (w > max_w) max_w = w;
        }
    const double avg_w = eff / (double)tiles_per_batch;
    i64 smax = ksteps / 4;
    if (smax > 64) smax = 64;
    const i64 part_elems = batch * n * ((m + 1) & ~(i64)1);
    while (smax > 1 && smax * part_elems * 8 > ((i64)1 << 30)) --smax; // partial workspace <= 1 GB
    double best_cost = -1.0;
    for (i64 s = 1; s <= smax; ++s) {
        const i64 rounds = rb_cdiv(tiles * s, num_sms);
        const double busiest = rounds == 1 ? max_w : (double)rounds * avg_w;
        // + the partial round trip through HBM (s writes + s reads of every tile, ~1/200 step per tile each)
        const double cost = busiest * (double)(rb_cdiv(ksteps, s) + 2) + (s > 1 ? (double)(2 * s * tiles) / 200.0 : 0.0);
        if (best_cost < 0.0 || cost < best_cost) { best_cost = cost; splits = s; }
    }
    // what the kernel will actually run: k per split is a whole number of 32-deep steps
    const i64 kper = rb_cdiv(ksteps, splits) * BK;
    return rb_cdiv(k, kper);
}


// ---- stream-K planning (host) ---------------------------------------------------------------------------------------------
// Cost of one 32-deep k step of tile (tm, tn) in 1/64 of a full tile step -- the same weights the split-K model uses (ragged tiles
// are worked on at 16 x 16 block granularity; diagonal tiles of triangular products as a list of their blocks; measured floor 0.22).
int tile_weight64(i64 tm, i64 tn, i64 tiles_m, i64 tiles_n, i64 m, i64 n, int tri)
{
    const i64 me = m - (tiles_m - 1) * BM, ne = n - (tiles_n - 1) * BN;
    const int mb = tm == tiles_m - 1 ? (int)((me + 15) >> 4) : 8, nb = tn == tiles_n - 1 ? (int)((ne + 15) >> 4) : 8;
    double w;
    if (tri && tm == tn) w = (double)((mb * (mb + 1) / 2 + 7) / 8) / 8.0;
    else {
        int best = 1 << 30; // blocks of the busiest warp under the best warp grid (<= 2 x 4 per warp)
        for (int lg = 3; lg >= 0; --lg) {
            const int r = (mb + (1 << lg) - 1) >> lg, c = (nb + (8 >> lg) - 1) >> (3 - lg);
            if (r <= 2 && c <= 4 && r * c < best) best = r * c;
        }
        w = (double)best / 8.0;
    }
    if (w < 0.22) w = 0.22;
    int wi = (int)(w * 64.0 + 0.5);
    return wi > 64 ? 64 : wi;
}

// tile index (within a batch) -> (tm, tn): the host twin of decode_tile
void host_decode_tile(i64 t, i64 tiles_m, i64 tiles_n, int tri, i64 &tm, i64 &tn)
{
    if (tri == 0) {
        const i64 per_group = GROUP_M * tiles_n, grp = t / per_group, first_m = grp * GROUP_M;
        const i64 gsz = tiles_m - first_m < GROUP_M ? tiles_m - first_m : GROUP_M, r = t - grp * per_group;
        tm = first_m + r % gsz; tn = r / gsz;
    } else {
        i64 hi = (i64)((std::sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
        while (hi * (hi + 1) / 2 > t) --hi;
        while ((hi + 1) * (hi + 2) / 2 <= t) ++hi;
        const i64 lo = t - hi * (hi + 1) / 2;
        if (tri == 1) { tm = lo; tn = hi; } else { tm = hi; tn = lo; }
    }
}

struct GemmPlan {
    i64 splits = 1;          // uniform split-K count (stream_k: the largest piece count)
    bool stream_k = false;
    int sk_grid = 0;
    SkTables sk;
    double cost_split = 0.0, cost_stream_k = 0.0; // estimated makespans in full-tile k steps (diagnostics / tests)
};

// Makespan of the uniform split under the kernel's dynamic scheduler (list scheduling in work-item order: the first G items are taken
// statically, every CTA then draws the next item when it runs dry), in full-tile k steps.  The closed-form model of plan_splits picks
// the split count; this is what it really costs, and what stream-K has to beat.
double simulate_split_makespan(const std::vector<int> &w64, i64 tiles_per_batch, i64 batch, i64 splits, i64 ksteps, int g)
{
    const i64 steps = rb_cdiv(ksteps, splits);
    std::vector<double> heap((size_t)g, 0.0); // min-heap of CTA finish times
    auto sift = [&](size_t i) {
        for (;;) {
            size_t l = 2 * i + 1, r = l + 1, mn = i;
            if (l < heap.size() && heap[l] < heap[mn]) mn = l;
            if (r < heap.size() && heap[r] < heap[mn]) mn = r;
            if (mn == i) return;
            std::swap(heap[i], heap[mn]); i = mn;
        }
    };
    double makespan = 0.0;
    for (i64 b = 0; b < batch; ++b)
        for (i64 t = 0; t < tiles_per_batch; ++t)
            for (i64 sp = 0; sp < splits; ++sp) {
                i64 st = ksteps - sp * steps; if (st > steps) st = steps; if (st <= 0) continue;
                heap[0] += (double)w64[(size_t)t] / 64.0 * (double)(st + 2);
                if (heap[0] > makespan) makespan = heap[0];
                sift(0);
            }
    return makespan;
}

// Build the stream-K tables for `tiles` tiles of ksteps steps each on at most num_sms CTAs; false when the shape does not qualify.
bool build_stream_k(GemmPlan &pl, const std::vector<int> &w64, i64 tiles_per_batch, i64 batch, i64 ksteps, int num_sms)
{
    const i64 tiles = tiles_per_batch * batch;
    if (tiles > SK_MAX_TILES || num_sms > SK_MAX_CTAS || ksteps >= (1LL << 31)) return false;
    i64 total = 0;
    for (i64 t = 0; t < tiles_per_batch; ++t) total += (i64)w64[(size_t)t] * ksteps;
    total *= batch;
    i64 g = num_sms;
    if (total / 64 < g) g = total / 64; // every CTA at least one full-tile step: no empty ranges
    if (g < 1) return false;
    // CTA c starts at the first step whose cost interval begins at or after floor(c total / g)
    i64 t = 0, wt0 = 0; // current tile and the cost at its first step
    for (i64 c = 0; c <= g; ++c) {
        const i64 bc = c == g ? total : (i64)((__int128)c * total / g);
        while (t < tiles && wt0 + (i64)w64[(size_t)(t % tiles_per_batch)] * ksteps <= bc) { wt0 += (i64)w64[(size_t)(t % tiles_per_batch)] * ksteps; ++t; }
        i64 st = 0;
        if (t < tiles) {
            const i64 w = w64[(size_t)(t % tiles_per_batch)];
            st = (bc - wt0 + w - 1) / w;
            if (st >= ksteps) { wt0 += w * ksteps; ++t; st = 0; }
        }
        pl.sk.cta_tile[c] = (unsigned short)t; pl.sk.cta_step[c] = (unsigned)st;
    }
    if (pl.sk.cta_tile[0] != 0 || pl.sk.cta_step[0] != 0 || pl.sk.cta_tile[g] != tiles || pl.sk.cta_step[g] != 0) return false;
    // pieces per tile; the busiest CTA
    i64 pmax = 1;
    double busiest = 0.0;
    for (i64 u = 0; u < tiles; ++u) { pl.sk.tile_first[u] = 0; pl.sk.tile_pieces[u] = 0; }
    for (i64 c = 0; c < g; ++c) {
        i64 ct = pl.sk.cta_tile[c], cs = pl.sk.cta_step[c];
        const i64 et = pl.sk.cta_tile[c + 1], es = pl.sk.cta_step[c + 1];
        if (!(ct < et || (ct == et && cs < es))) return false; // an empty range would break the piece numbering
        double cost = 0.0;
        while (ct < et || (ct == et && cs < es)) {
            const i64 s1 = ct == et ? es : ksteps;
            if (pl.sk.tile_pieces[ct] == 0) pl.sk.tile_first[ct] = (unsigned char)c;
            if (++pl.sk.tile_pieces[ct] == 0) return false; // > 255 pieces
            cost += (double)w64[(size_t)(ct % tiles_per_batch)] / 64.0 * (double)(s1 - cs) + 1.5; // + fill / partial store per piece
            ++ct; cs = 0;
        }
        if (cost > busiest) busiest = cost;
    }
    for (i64 u = 0; u < tiles; ++u) {
        if (pl.sk.tile_pieces[u] == 0) return false;
        if (pl.sk.tile_pieces[u] > pmax) pmax = pl.sk.tile_pieces[u];
    }
    pl.sk_grid = (int)g;
    pl.splits = pmax;
    pl.cost_stream_k = busiest + 1.5 + (double)(4 * (tiles < g ? tiles : g)) / 200.0 + 0.05 * (double)pmax; // + the fix-up pass (a chain of pmax loads per element)
    return true;
}

// Decide between the uniform split (count `splits`, already chosen by plan_splits) and stream-K for one shape.
void make_plan(GemmPlan &pl, i64 m, i64 n, i64 k, i64 batch, int tri, int num_sms, i64 splits, i64 tiles_m, i64 tiles_n, i64 tiles_per_batch)
{
    static const int sk_on = [] { const char *e = getenv("REST_B200_STREAMK"); return e ? atoi(e) : 1; }();
    const i64 ksteps = rb_cdiv(k, BK);
    std::vector<int> w64((size_t)tiles_per_batch);
    for (i64 t = 0; t < tiles_per_batch; ++t) {
        i64 tm, tn;
        host_decode_tile(t, tiles_m, tiles_n, tri, tm, tn);
        w64[(size_t)t] = tile_weight64(tm, tn, tiles_m, tiles_n, m, n, tri);
    }
    pl.splits = splits; pl.stream_k = false;
    const i64 tiles = tiles_per_batch * batch, items = tiles * splits;
    const int g = (int)(items < num_sms ? items : num_sms);
    pl.cost_split = simulate_split_makespan(w64, tiles_per_batch, batch, splits, ksteps, g) +
                    (splits > 1 ? 1.5 + (double)(2 * splits * tiles) / 200.0 + 0.05 * (double)splits : 0.0);
    GemmPlan sk;
    const i64 ldp = (m + 1) & ~(i64)1;
    bool have_sk = false;
    for (int div = 1; div <= 4 && sk_on; div *= 2) { // all SMs, or fewer CTAs with longer pieces when the fix-up chain would dominate
        GemmPlan cand;
        if (num_sms / div < 1 || !build_stream_k(cand, w64, tiles_per_batch, batch, ksteps, num_sms / div)) continue;
        if (!have_sk || cand.cost_stream_k < sk.cost_stream_k) { sk = cand; have_sk = true; }
    }
    if (have_sk && sk.splits * batch * n * ldp * 8 <= ((i64)1 << 30)) {
        pl.cost_stream_k = sk.cost_stream_k;
        if (sk.cost_stream_k < 0.97 * pl.cost_split) {
            const double cs = pl.cost_split;
            pl = sk; pl.cost_split = cs; pl.stream_k = true;
        }
    }
}

struct PlanKey { i64 m, n, k, batch; int tri, num_sms; i64 splits; };
const GemmPlan *cached_plan(i64 m, i64 n, i64 k, i64 batch, int tri, int num_sms, i64 splits, i64 tiles_m, i64 tiles_n, i64 tiles_per_batch)
{
    constexpr int SLOTS = 16;
    static thread_local PlanKey keys[SLOTS];
    static thread_local GemmPlan plans[SLOTS];
    static thread_local unsigned stamp[SLOTS], clock_ = 0;
    static thread_local bool used[SLOTS];
    ++clock_;
    int victim = 0;
    for (int i = 0; i < SLOTS; ++i) {
        if (used[i] && keys[i].m == m && keys[i].n == n && keys[i].k == k && keys[i].batch == batch && keys[i].tri == tri &&
            keys[i].num_sms == num_sms && keys[i].splits == splits) { stamp[i] = clock_; return &plans[i]; }
        if (!used[i]) victim = i;
        else if (used[victim] && stamp[i] < stamp[victim]) victim = i;
    }
    plans[victim] = GemmPlan();
    make_plan(plans[victim], m, n, k, batch, tri, num_sms, splits, tiles_m, tiles_n, tiles_per_batch);
    keys[victim] = PlanKey{m, n, k, batch, tri, num_sms, splits};
    used[victim] = true; stamp[victim] = clock_;
    return &plans[victim];
}

template <bool A_K, bool B_K, bool PACK>
int launch_tma(rb_ctx *ctx, const CUtensorMap &tmA, const CUtensorMap &tmB, const GemmParams &p, int grid)
{
    static bool attr_set[64] = {false}; // the opt-in shared-memory size is a per-device function attribute
    const int dev = ctx->device & 63;
    if (!attr_set[dev]) {
        RB_CUDA(cudaFuncSetAttribute(rb_gemm_tma_kernel<A_K, B_K, PACK>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr_set[dev] = true;
    }
    rb_gemm_tma_kernel<A_K, B_K, PACK><<<grid, NUM_THREADS, SMEM_BYTES, ctx->stream>>>(tmA, tmB, p);
    RB_LAUNCHED(ctx);
    return RB_OK;
}

} // namespace

extern "C" int64_t rb_gemm_plan_splits(int64_t m, int64_t n, int64_t k, int64_t batch, int tri, int num_sms)
{
    if (m <= 0 || n <= 0 || k <= 0 || batch <= 0) return 1;
    return plan_splits(m, n, k, batch, tri, num_sms);
}

// 1 when the (non-packed) product would run as stream-K; costs_out[0..1] (optional) = estimated makespans of the uniform split and of
// stream-K in full-tile k steps, costs_out[2] = CTAs, costs_out[3] = largest piece count
extern "C" int rb_gemm_plan_stream_k(int64_t m, int64_t n, int64_t k, int64_t batch, int tri, int num_sms, double *costs_out)
{
    if (costs_out) { costs_out[0] = costs_out[1] = costs_out[2] = costs_out[3] = 0.0; }
    if (m <= 0 || n <= 0 || k <= 0 || batch <= 0 || num_sms <= 0) return 0;
    if (tri && m != n) return 0;
    const i64 tiles_m = rb_cdiv(m, BM), tiles_n = rb_cdiv(n, BN);
    const i64 tiles_per_batch = tri ? tiles_m * (tiles_m + 1) / 2 : tiles_m * tiles_n, ksteps = rb_cdiv(k, BK);
    if (!(tiles_per_batch * batch < 4 * (i64)num_sms && ksteps >= 8)) return 0;
    i64 splits = plan_splits(m, n, k, batch, tri, num_sms);
    splits = rb_cdiv(k, rb_cdiv(ksteps, splits) * BK);
    GemmPlan pl;
    make_plan(pl, m, n, k, batch, tri, num_sms, splits, tiles_m, tiles_n, tiles_per_batch);
    if (costs_out) { costs_out[0] = pl.cost_split; costs_out[1] = pl.cost_stream_k; costs_out[2] = pl.sk_grid; costs_out[3] = (double)pl.splits; }
    return pl.stream_k ? 1 : 0;
}

// the stream-K tables of a shape, for the CPU tests of the partition (returns the CTA count, 0 when the shape does not qualify)
extern "C" int rb_gemm_stream_k_tables(int64_t m, int64_t n, int64_t k, int64_t batch, int tri, int num_sms, unsigned *cta_step,
                                       unsigned short *cta_tile, unsigned char *tile_first, unsigned char *tile_pieces)
{
    if (m <= 0 || n <= 0 || k <= 0 || batch <= 0 || num_sms <= 0 || (tri && m != n)) return 0;
    const i64 tiles_m = rb_cdiv(m, BM), tiles_n = rb_cdiv(n, BN);
    const i64 tiles_per_batch = tri ? tiles_m * (tiles_m + 1) / 2 : tiles_m * tiles_n, ksteps = rb_cdiv(k, BK);
    if (tiles_per_batch * batch > SK_MAX_TILES) return 0;
    std::vector<int> w64((size_t)tiles_per_batch);
    for (i64 t = 0; t < tiles_per_batch; ++t) {
        i64 tm, tn;
        host_decode_tile(t, tiles_m, tiles_n, tri, tm, tn);
        w64[(size_t)t] = tile_weight64(tm, tn, tiles_m, tiles_n, m, n, tri);
    }
    GemmPlan pl;
    if (!build_stream_k(pl, w64, tiles_per_batch, batch, ksteps, num_sms)) return 0;
    for (int c = 0; c <= pl.sk_grid; ++c) { cta_step[c] = pl.sk.cta_step[c]; cta_tile[c] = pl.sk.cta_tile[c]; }
    for (i64 t = 0; t < tiles_per_batch * batch; ++t) { tile_first[t] = pl.sk.tile_first[t]; tile_pieces[t] = pl.sk.tile_pieces[t]; }
    return pl.sk_grid;
}

int rb_gemm_core(rb_ctx *ctx, bool ta, bool tb, i64 m, i64 n, i64 k, double alpha, const double *a, i64 lda,
                 i64 stride_a, const double *b, i64 ldb, i64 stride_b, double beta, double *c, i64 ldc, i64 stride_c,
                 i64 batch, int tri)
{
    if (m <= 0 || n <= 0 || batch <= 0) return RB_OK;
    RB_REQUIRE(c != nullptr, "gemm: C is NULL");
    RB_REQUIRE(ldc >= m, "gemm: ldc (%lld) < m (%lld)", (long long)ldc, (long long)m);
    if (k <= 0 || alpha == 0.0) {
        if (beta == 1.0) return RB_OK;
        i64 total = m * n * batch;
        i64 blocks = rb_cdiv(total, 256);
        i64 cap = (i64)ctx->num_sms * 16;
        if (blocks > cap) blocks = cap;
        rb_scale_c_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(c, m, n, ldc, stride_c, batch, beta, tri);
        RB_LAUNCHED(ctx);
        return RB_OK;
    }
    RB_REQUIRE(a && b, "gemm: NULL operand");
    RB_REQUIRE(lda >= (ta ? k : m), "gemm: lda (%lld) too small", (long long)lda);
    RB_REQUIRE(ldb >= (tb ? n : k), "gemm: ldb (%lld) too small", (long long)ldb);
    if (tri) RB_REQUIRE(m == n, "gemm: triangular output needs m == n");

    const bool a_k = ta, b_k = !tb; // K-major operands
    bool a_ok = tma_eligible(ctx, a, lda, stride_a, batch, a_k ? k : m, a_k ? m : k);
    bool b_ok = tma_eligible(ctx, b, ldb, stride_b, batch, b_k ? k : n, b_k ? n : k);
    // An operand that only fails on alignment (odd leading dimension / batch stride, 8-byte-aligned base) is copied
    // into a dense even-pitch workspace block and the TMA kernel runs on the copy: one HBM pass over the operand
    // instead of the 6x slower plain-load kernel (2049^3: 6 vs ~33 TFLOP/s).  The pad element of an odd extent lies
    // outside the tensor map's bounds, so it is never read.
    if (ctx->gemm_path == 0 && ctx->encode_tiled && k < (1LL << 31) && (!a_ok || !b_ok) && m * n * k >= (1LL << 21)) {
        const bool same = (a == b && lda == ldb && stride_a == stride_b && a_k == b_k && m == n); // SYRK: one copy serves both
        auto extents = [&](bool is_a, i64 &d0, i64 &d1, i64 &bat) {
            const bool km = is_a ? a_k : b_k;
            const i64 rows = is_a ? m : n;
            d0 = km ? k : rows; d1 = km ? rows : k;
            bat = (batch > 1 && (is_a ? stride_a : stride_b) != 0) ? batch : 1;
        };
        i64 a0, a1, ab, b0, b1, bb;
        extents(true, a0, a1, ab); extents(false, b0, b1, bb);
        const i64 la = a0 + (a0 & 1), lb = b0 + (b0 & 1);
        const i64 ea = a_ok ? 0 : la * a1 * ab, eb = (b_ok || same) ? 0 : lb * b1 * bb;
        const bool dims_ok = a0 < (1LL << 31) && a1 < (1LL << 31) && b0 < (1LL << 31) && b1 < (1LL << 31) && batch < (1LL << 31);
        if (dims_ok && (ea + eb) * 8 <= ((i64)4 << 30)) {
            void *ws;
            RB_TRY(rb_ws_reserve(ctx, 2, (ea + eb) * 8 + 32, &ws));
            double *pa = (double *)ws, *pb = pa + ea;
            if (!a_ok) {
                RB_TRY(rb_copy3d(ctx, a, 0, 1, lda, stride_a, pa, 0, 1, la, la * a1, a0, a1, ab));
                a = pa; lda = la; if (ab > 1) stride_a = la * a1;
                a_ok = true;
                if (same) { b = pa; ldb = la; stride_b = stride_a; b_ok = true; }
            }
            if (!b_ok) {
                RB_TRY(rb_copy3d(ctx, b, 0, 1, ldb, stride_b, pb, 0, 1, lb, lb * b1, b0, b1, bb));
                b = pb; ldb = lb; if (bb > 1) stride_b = lb * b1;
                b_ok = true;
            }
        }
    }
    bool use_tma = ctx->gemm_path == 0 && a_ok && b_ok;
    if (use_tma) {
        CUtensorMap tmA, tmB;
        const bool a_batched = batch > 1 && stride_a != 0, b_batched = batch > 1 && stride_b != 0;
        RB_TRY(encode_map(ctx, &tmA, a, a_k, m, k, lda, stride_a, a_batched ? batch : 1));
        RB_TRY(encode_map(ctx, &tmB, b, b_k, n, k, ldb, stride_b, b_batched ? batch : 1));
        GemmParams p{};
        p.m = m; p.n = n; p.k = k; p.batch = batch;
        const bool packed = pack_m_applies(a_k, a_batched, b_batched, m, batch, tri);
        p.pack_m = packed ? 1 : 0;
        p.bpb = rb_cdiv(m, 16); p.total_blocks = p.bpb * batch;
        p.tiles_m = packed ? rb_cdiv(p.total_blocks, BM / 16) : rb_cdiv(m, BM); p.tiles_n = rb_cdiv(n, BN);
        p.tiles_per_batch = tri ? p.tiles_m * (p.tiles_m + 1) / 2 : p.tiles_m * p.tiles_n;
        i64 tiles = p.tiles_per_batch * (packed ? 1 : batch);
        const i64 ksteps = rb_cdiv(k, BK);
        i64 splits = plan_splits(m, n, k, batch, tri, ctx->num_sms, packed);
        i64 kper = rb_cdiv(ksteps, splits) * BK;
        splits = rb_cdiv(k, kper);
        // stream-K instead of the uniform split where the tiles do not fill the chip (same trigger as split-K) and the simulated
        // makespan says so; plans are cached per shape (an SCF loop repeats a handful of shapes)
        p.stream_k = 0; p.sk_ksteps = ksteps;
        int sk_grid = 0;
        if (!packed && tiles < 4 * (i64)ctx->num_sms && ksteps >= 8) {
            const GemmPlan *pl = cached_plan(m, n, k, batch, tri, ctx->num_sms, splits, p.tiles_m, p.tiles_n, p.tiles_per_batch);
            if (pl && pl->stream_k) {
                p.stream_k = 1; p.sk = pl->sk; sk_grid = pl->sk_grid;
                splits = pl->splits; kper = ksteps * BK;
            }
        }
        const bool stream_k = p.stream_k != 0;
        p.splits = splits; p.kper = kper;
        p.total_items = tiles * splits;
        p.alpha = alpha; p.beta = beta; p.c = c; p.ldc = ldc; p.stride_c = stride_c; p.tri = tri;
        p.partial = nullptr; p.ldp = (m + 1) & ~(i64)1;
        p.a_batched = a_batched ? 1 : 0; p.b_batched = b_batched ? 1 : 0;
        // same matrix on both sides of a triangular product (SYRK): diagonal tiles need one operand tile only
        p.same_ab = (tri != 0 && a == b && lda == ldb && a_k == b_k && (batch <= 1 || stride_a == stride_b)) ? 1 : 0;
        // one of 64 self-re-arming scheduler slots per launch: launches of one context may overlap (the caller can
        // move the context between streams) without sharing a work counter
        p.sched = ctx->sched + 2 * (ctx->sched_next++ & 63);
        if (splits > 1) {
            void *ws;
            RB_TRY(rb_ws_reserve(ctx, 1, splits * batch * n * p.ldp * 8, &ws));
            p.partial = (double *)ws;
        }
        int grid = (int)(p.total_items < ctx->num_sms ? p.total_items : ctx->num_sms);
        if (stream_k) grid = sk_grid;
        if (a_k && b_k) RB_TRY((launch_tma<true, true, false>(ctx, tmA, tmB, p, grid)));
        else if (a_k && !b_k) RB_TRY((launch_tma<true, false, false>(ctx, tmA, tmB, p, grid)));
        else if (!a_k && b_k) {
            if (packed) RB_TRY((launch_tma<false, true, true>(ctx, tmA, tmB, p, grid)));
            else RB_TRY((launch_tma<false, true, false>(ctx, tmA, tmB, p, grid)));
        } else {
            if (packed) RB_TRY((launch_tma<false, false, true>(ctx, tmA, tmB, p, grid)));
            else RB_TRY((launch_tma<false, false, false>(ctx, tmA, tmB, p, grid)));
        }
        if (splits > 1) {
            // (row pairs, columns, batches); the pad row of an odd m (ldp = m + 1) is never read
            i64 bx = rb_cdiv(rb_cdiv(m, 2), 256);
            if (bx > 64) bx = 64;
            const i64 by = n < 65535 ? n : 65535, bz = batch < 64 ? batch : 64;
            dim3 rgrid((unsigned)bx, (unsigned)by, (unsigned)bz);
            if (stream_k)
                rb_partial_reduce_kernel<true><<<rgrid, 256, 0, ctx->stream>>>(p.partial, p.ldp, (int)splits, m, n, batch, alpha, beta, c, ldc, stride_c,
                                                                              tri, (int)p.tiles_m, (int)p.tiles_n, (int)p.tiles_per_batch, p.sk);
            else
                rb_partial_reduce_kernel<false><<<rgrid, 256, 0, ctx->stream>>>(p.partial, p.ldp, (int)splits, m, n, batch, alpha, beta, c, ldc, stride_c,
                                                                               tri, (int)p.tiles_m, (int)p.tiles_n, (int)p.tiles_per_batch, p.sk);
            RB_LAUNCHED(ctx);
        }
        return RB_OK;
    }
    // generic path
    i64 tiles_m = rb_cdiv(m, GB), tiles_n = rb_cdiv(n, GB);
    RB_REQUIRE(tiles_m * tiles_n < 2147483647LL && batch <= 65535, "gemm(generic): problem too large for this path");
    dim3 grid((unsigned)(tiles_m * tiles_n), (unsigned)batch);
    rb_gemm_generic_kernel<<<grid, 128, 0, ctx->stream>>>(ta, tb, m, n, k, alpha, a, lda, stride_a, b, ldb, stride_b,
                                                          beta, c, ldc, stride_c, tiles_m, tiles_n, tri);
    RB_LAUNCHED(ctx);
    return RB_OK;
}
