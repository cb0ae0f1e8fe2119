// rb_dpj.cu -- d_P and J in ONE pass over ri3ao (SURVEY 8a rows a7 + a8; the reference makes two passes: a dgemv 'T' for
// d[P] = sum_ij A[ij,P] D[ij] and a dgemv 'N' for J[ij] = sum_P A[ij,P] d[P]).
//
// Both products read every element of ri3ao once and are bound by that read, but J needs d_P of a slab only after the WHOLE slab has been
// dotted with D.  One persistent, cooperative kernel (one CTA per SM) does both from a single HBM read:
//   * CTA c owns the same range of ij (L = nb^2 / #CTAs elements) in EVERY slab; its D[ij] and its J[ij] live in registers for the
//     whole kernel, so no cross-CTA reduction of J is ever needed;
//   * the CTA's L-element run of each slab streams into a shared-memory ring with 16-byte LDGSTS copies (cp.async.cg: global -> shared
//     without registers); every thread copies exactly the units it reads back, so the ring is synchronised by per-thread commit groups
//     alone; ~100-190 KB in flight per SM.  (1-D bulk copies -- cp.async.bulk + mbarriers -- were measured first: they saturate at
//     22.6 GB/s per SM = 3.3 TB/s, half of what the pass needs.)
//   * stage 1 (as soon as a block of S slabs has landed): partial dots of the block with D, one value per slab and CTA, stored into a
//     [block][slab][CTA] array that the host pre-fills with a sentinel (all-ones bit pattern, a NaN no arithmetic produces): the value
//     IS the message -- no counters, no fences, no atomics;
//   * stage 2 (LAG blocks later, so that the other CTAs' values have normally long arrived): the loads of the #CTAs partials of a slab are
//     issued BEFORE stage 1 of the current block and checked after it (a sentinel means "not there yet": poll), so the L2 round trip
//     hides behind the dot products; they are added in CTA order -> d_P (bitwise the same in every CTA), then
//     J[ij] += d_P * A[ij,P] from the copy that is STILL in shared memory, and the ring slot goes back to the loader.
// HBM traffic: nb^2 * nx * 8 bytes once (plus kilobytes of partials) instead of twice.  Deterministic: fixed thread / warp / CTA order.
// 16 compute warps + 2 exchange warps per CTA: the exchange warps publish the CTA's partial dots and gather everybody's, one block ahead
// of the compute warps, so the grid-wide exchange is off the critical path.  What bounds the kernel is shared memory: loads in flight
// (2.4 us of latency x 44 GB/s per SM = 106 KB) + the block being worked on + the blocks waiting for their d_P have to fit 200 KB, which
// leaves one to two blocks in flight -> 4.85 TB/s of the single read = 1.34-1.46x the two passes for nb = 600-900.
// Shapes the ring cannot hold (nb > ~1100), odd nb^2 alignment, short runs (nb < ~550) or small tensors use the two GEMV kernels.
#include "rb_common.cuh"

namespace {

constexpr int DPJ_THREADS = 512;
constexpr int DPJ_WARPS = DPJ_THREADS / 32;   // compute warps
constexpr int DPJ_XWARPS = 2;                  // exchange warps (publish the CTA's partial dots, gather everybody's -> d_P)
constexpr int DPJ_BLOCK = DPJ_THREADS + 32 * DPJ_XWARPS;
constexpr int DPJ_MAX_S = 8;
constexpr int DPJ_MAX_RING = 10;
constexpr i64 DPJ_RING_BYTES = 220 * 1024;

struct DpjParams {
    const double *a;
    i64 slab, nx;       // nb * nb, number of slabs
    const double *dm;   // D [nb * nb]
    double *d, *j;      // d [nx], J [nb * nb]
    i64 L;              // elements of a slab per CTA (multiple of 4)
    int S, R, LAG;      // slabs per block, ring depth in blocks, blocks between stage 1 and stage 2
    i64 nblocks;
    double *partial;    // [nblocks][S][gridDim.x], every element the sentinel on entry
    int debug;          // 1: tools/prof_dpj_sweep.py only -- skip the cross-CTA gather (WRONG results; isolates the cost of the exchange)
};

constexpr int DPJ_NOT_RESIDENT = -77; // internal status of dpj_launch: cooperative launch impossible here
constexpr long long DPJ_SENTINEL = -1LL; // all ones: a NaN payload that FMA / add never produce

__device__ __forceinline__ uint32_t dpj_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// 16 bytes global -> shared without passing through registers (LDGSTS), L2 only; completion is tracked per thread in commit groups
__device__ __forceinline__ void dpj_cp16(uint32_t dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ double2 dpj_lds128(uint32_t addr)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void dpj_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void dpj_wait_pending(int n) // wait until at most n of this thread's newest groups are still in flight
{
    switch (n) {
    case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
    case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
    case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
    case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
    case 4: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
    case 5: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
    case 6: asm volatile("cp.async.wait_group 6;" ::: "memory"); break;
    case 7: asm volatile("cp.async.wait_group 7;" ::: "memory"); break;
    default: asm volatile("cp.async.wait_group 8;" ::: "memory"); break;
    }
}
__device__ __forceinline__ double dpj_ld_gpu(const double *p) // coherent at GPU scope: never served from this SM's L1
{
    double v;
    asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void dpj_st_gpu(double *p, double v)
{
    asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ double dpj_warp_sum(double v) // fixed butterfly: every lane ends with the same bits
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// K = 16-byte units per thread that cover the CTA's run (L <= K * 2 * DPJ_THREADS), ST = slabs per block (compile time: the ST dot
// products and their butterflies are interleaved, not chained).
template <int K, int ST>
__global__ void __launch_bounds__(DPJ_BLOCK, 1) rb_ri_dp_j_kernel(const DpjParams p)
{
    extern __shared__ __align__(128) unsigned char dpj_smem[];
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int Gi = (int)gridDim.x, cta = (int)blockIdx.x;
    const i64 e0 = (i64)cta * p.L;
    const int len = (int)(p.slab - e0 < p.L ? (p.slab - e0 > 0 ? p.slab - e0 : 0) : p.L); // multiple of 4 (slab and L are)
    double *ring = reinterpret_cast<double *>(dpj_smem);
    const int R = p.R, LAG = p.LAG, nblocks = (int)p.nblocks;
    const uint32_t run_bytes = (uint32_t)p.L * 8u, block_bytes = run_bytes * ST;
    double *red = ring + (size_t)R * ST * p.L;       // [2][ST][DPJ_WARPS]  (by iteration parity)
    double *dsm = red + 2 * ST * DPJ_WARPS;          // [2][ST]             (by block parity)
    const uint32_t ring_s = dpj_smem_u32(ring);

    // the ring starts as zeros: slabs past the end of the last block are read (with weight zero) like any other
    for (uint32_t o = (uint32_t)t * 16u; o < (uint32_t)R * block_bytes; o += DPJ_BLOCK * 16u)
        *reinterpret_cast<double2 *>(dpj_smem + o) = make_double2(0.0, 0.0);
    __syncthreads();

    constexpr int GV = 5; // partials per lane: up to 160 CTAs
    const bool xrole = warp >= DPJ_WARPS;
    // ===================== exchange warps: off the compute warps' critical path =====================
    // (both roles run ONE loop; they are coupled by two named barriers, see the end of the loop)
    {
        // iteration `it`: (a) when the compute warps have left their partial dots of block `it` in red[it & 1] (named barrier 1; barrier 2, all 576 threads,
        // closes every iteration -- named barriers with explicit counts, because the two roles meet from different instructions), add
        // them over the warps and publish the CTA's value; (b) gather block g = it - LAG + 1 from all CTAs -> dsm[g & 1], which the
        // compute warps use in the NEXT iteration.  The gather's loads are issued first, so their L2 round trip overlaps (a).
    }
    const int xw = warp - DPJ_WARPS;
    double gv[GV];
#pragma unroll
    for (int i = 0; i < GV; ++i) gv[i] = 0.0;
    auto exchange_iteration = [&](int it) {
        {
            const int g = it - LAG + 1;
            const bool g_ok = g >= 0 && g < nblocks;
            const int s_ng = g_ok ? (int)(p.nx - (i64)g * ST < ST ? p.nx - (i64)g * ST : ST) : 0;
            if (p.debug != 1 && xw < s_ng) {
                const double *row = p.partial + (size_t)(g * ST + xw) * Gi;
#pragma unroll
                for (int i = 0; i < GV; ++i) gv[i] = (lane + 32 * i < Gi) ? dpj_ld_gpu(row + lane + 32 * i) : 0.0;
            }
            __syncwarp();
            asm volatile("bar.sync 1, %0;" ::"n"(DPJ_BLOCK) : "memory"); // every iteration, so that neither role runs a barrier phase ahead
            if (it < nblocks) {
                const int s_n = (int)(p.nx - (i64)it * ST < ST ? p.nx - (i64)it * ST : ST);
                for (int s = xw; s < s_n; s += DPJ_XWARPS) {
                    double v = lane < DPJ_WARPS ? red[((it & 1) * ST + s) * DPJ_WARPS + lane] : 0.0;
                    v = dpj_warp_sum(v); // fixed butterfly over the 16 compute warps
                    if (lane == 0) dpj_st_gpu(p.partial + (size_t)(it * ST + s) * Gi + cta, v);
                }
            }
            for (int s = xw; s < ST; s += DPJ_XWARPS) {
                double v = 0.0; // slabs past the end of the last block get weight zero
                if (s < s_ng) {
                    if (p.debug == 1) v = 1.0;
                    else {
                        const double *row = p.partial + (size_t)(g * ST + s) * Gi;
                        if (s != xw) {
#pragma unroll
                            for (int i = 0; i < GV; ++i) gv[i] = (lane + 32 * i < Gi) ? dpj_ld_gpu(row + lane + 32 * i) : 0.0;
                        }
                        const long long t0 = clock64();
                        for (;;) { // a sentinel = that CTA has not published yet
                            bool missing = false;
#pragma unroll
                            for (int i = 0; i < GV; ++i)
                                if (__double_as_longlong(gv[i]) == DPJ_SENTINEL) { gv[i] = dpj_ld_gpu(row + lane + 32 * i); missing = true; }
                            if (!__any_sync(0xffffffffu, missing)) break;
                            if (clock64() - t0 > 8000000000LL) __trap(); // a lost CTA must not hang the device: fail loudly instead
                        }
#pragma unroll
                        for (int i = 0; i < GV; ++i) v += gv[i]; // lane: CTAs lane, lane + 32, ... in order; then the fixed butterfly
                        v = dpj_warp_sum(v);
                        if (lane == 0 && cta == 0) p.d[g * ST + s] = v;
                    }
                }
                if (g_ok && lane == 0) dsm[(g & 1) * ST + s] = v;
            }
            if (it + 1 < nblocks + LAG) { // this iteration's dsm is complete: the compute warps wait for it in their NEXT iteration
                __syncwarp();
                asm volatile("bar.arrive 3, %0;" ::"n"(DPJ_BLOCK) : "memory");
            }
        }
    };

    // ===================== compute warps =====================
    // Per-thread geometry: unit k of this thread is doubles [q_k, q_k + 2) of the run.  Units past the end of the run alias the thread's
    // first unit with weight zero, so the loops below carry no predicates; a thread without any unit sits the arithmetic out.
    const bool tv = !xrole && 2 * t < len;
    uint32_t off[K];
    bool ok[K];
    double2 dreg[K], jreg[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int q = 2 * (t + k * DPJ_THREADS);
        ok[k] = !xrole && q < len;
        off[k] = (uint32_t)(ok[k] ? q : 2 * t) * 8u;
        dreg[k] = ok[k] ? *reinterpret_cast<const double2 *>(p.dm + e0 + q) : make_double2(0.0, 0.0);
        jreg[k] = make_double2(0.0, 0.0);
    }

    // Every thread copies exactly the 16-byte units it reads back itself, so the ring needs no barrier of its own: a thread waits for
    // its own commit groups.  One group per block, committed by every thread whether or not it copied.
    const double *src_next = p.a + e0; // first slab of the next block to be issued
    int left_issue = (int)p.nx;
    auto issue = [&](int slot) {
        const uint32_t dst = ring_s + (uint32_t)slot * block_bytes;
#pragma unroll
        for (int s = 0; s < ST; ++s) {
            if (s < left_issue) {
#pragma unroll
                for (int k = 0; k < K; ++k)
                    if (ok[k]) dpj_cp16(dst + s * run_bytes + off[k], src_next + (i64)s * p.slab + (off[k] >> 3));
            }
        }
        dpj_commit();
        src_next += (i64)ST * p.slab;
        left_issue -= ST;
    };
    if (!xrole)
        for (int b = 0; b < R; ++b) { // prologue: R groups (empty ones past the end keep the group arithmetic uniform)
            if (b < nblocks) issue(b);
            else dpj_commit();
        }

    int slot1 = 0, slot2 = 0;
    for (int it = 0; it < nblocks + LAG; ++it) {
        if (xrole) {
            exchange_iteration(it);
        } else {
        const int bb = it - LAG;
        if (it < nblocks) { // ---- stage 1: the ST partial dots of block `it`, interleaved
            // groups committed so far: R + max(0, it - LAG); block `it` is group `it`
            dpj_wait_pending(it >= LAG ? R - LAG - 1 : R - it - 1);
            const uint32_t base = ring_s + (uint32_t)slot1 * block_bytes;
            if (++slot1 == R) slot1 = 0;
            double acc[ST];
#pragma unroll
            for (int s = 0; s < ST; ++s) acc[s] = 0.0;
            if (tv) {
#pragma unroll
                for (int k = 0; k < K; ++k) {
#pragma unroll
                    for (int s = 0; s < ST; ++s) {
                        const double2 v = dpj_lds128(base + s * run_bytes + off[k]);
                        acc[s] += dreg[k].x * v.x;
                        acc[s] += dreg[k].y * v.y;
                    }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                for (int s = 0; s < ST; ++s) acc[s] += __shfl_xor_sync(0xffffffffu, acc[s], o);
            }
            if (lane == 0) {
#pragma unroll
                for (int s = 0; s < ST; ++s) red[((it & 1) * ST + s) * DPJ_WARPS + warp] = acc[s];
            }
        }
        // Hand-shake with the exchange warps, in this order: first wait until they are done with the PREVIOUS iteration (normally long
        // ago: they had this iteration's stage 1 and the last stage 2 to do it), only then hand them red[it & 1] -- so that no barrier ever
        // receives the arrivals of two iterations.
        __syncwarp();
        if (it >= 1) asm volatile("bar.sync 3, %0;" ::"n"(DPJ_BLOCK) : "memory");
        asm volatile("bar.arrive 1, %0;" ::"n"(DPJ_BLOCK) : "memory");
        if (bb >= 0) { // ---- stage 2: J += d_P * A from the copy still in the ring (d_P gathered during the previous iteration)
            const uint32_t base = ring_s + (uint32_t)slot2 * block_bytes;
            const int slot = slot2;
            if (++slot2 == R) slot2 = 0;
            if (tv) {
                double ds[ST];
#pragma unroll
                for (int s = 0; s < ST; ++s) ds[s] = dsm[(bb & 1) * ST + s];
#pragma unroll
                for (int k = 0; k < K; ++k) {
#pragma unroll
                    for (int s = 0; s < ST; ++s) {
                        const double2 v = dpj_lds128(base + s * run_bytes + off[k]);
                        jreg[k].x += ds[s] * v.x;
                        jreg[k].y += ds[s] * v.y;
                    }
                }
            }
            // the slot is private per thread (each thread reads only what it copied), so it can be refilled right away
            if (bb + R < nblocks) issue(slot);
            else dpj_commit();
        }
        } // compute role
        // No block-wide barrier per iteration: the two roles are coupled by the named barriers alone -- 1: red[it & 1] ready (compute
        // arrives, exchange waits), 3: exchange iteration done (exchange arrives, compute waits in its next iteration, before arriving
        // on 1 again).  red[it & 1] is rewritten in stage 1 of iteration it + 2, after the compute warps passed barrier 3 in iteration
        // it + 1, i.e. after the exchange warps finished iteration it; dsm[g & 1] is rewritten in exchange iteration it + 2, which starts
        // behind barrier 1 of it + 2, i.e. after the compute warps' stage 2 of iteration it + 1 that read it.
    }
    if (xrole) return;
#pragma unroll
    for (int k = 0; k < K; ++k)
        if (ok[k]) *reinterpret_cast<double2 *>(p.j + e0 + (off[k] >> 3)) = jreg[k];
}

template <int K, int ST>
int dpj_launch(rb_ctx *ctx, const DpjParams &p, size_t smem)
{
    static int configured_for = -1; // per process and device: the attribute belongs to the function on that device
    if (configured_for != ctx->device) {
        RB_CUDA(cudaFuncSetAttribute(rb_ri_dp_j_kernel<K, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(DPJ_RING_BYTES + 4096)));
        configured_for = ctx->device;
    }
    void *args[] = {(void *)&p};
    const cudaError_t e = cudaLaunchCooperativeKernel((const void *)rb_ri_dp_j_kernel<K, ST>, dim3((unsigned)ctx->num_sms), dim3(DPJ_BLOCK),
                                                      args, smem, ctx->stream);
    if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorNotSupported) {
        // not every SM is available to this process (MPS share, MIG slice, ...): the grid-wide exchange needs all CTAs resident, so
        // the caller runs the two GEMV passes instead (still on the GPU)
        (void)cudaGetLastError();
        return DPJ_NOT_RESIDENT;
    }
    RB_CUDA(e);
    RB_LAUNCHED(ctx);
    return RB_OK;
}

template <int ST>
int dpj_dispatch_k(rb_ctx *ctx, const DpjParams &p, size_t smem, int kk)
{
    switch (kk) {
    case 1: return dpj_launch<1, ST>(ctx, p, smem);
    case 2: return dpj_launch<2, ST>(ctx, p, smem);
    case 3: return dpj_launch<3, ST>(ctx, p, smem);
    case 4: return dpj_launch<4, ST>(ctx, p, smem);
    case 5: case 6: return dpj_launch<6, ST>(ctx, p, smem);
    default: return dpj_launch<8, ST>(ctx, p, smem);
    }
}

} // namespace

// d[P] = sum_ij ri3ao[ij, P] dm[ij]  and  j[ij] = sum_P ri3ao[ij, P] d[P]  (both overwritten) from one read of ri3ao.
extern "C" int rb_ri_dp_j(rb_ctx *ctx, const double *ri3ao, const double *dm, double *d, double *j, int nb, int nx)
{
    RB_REQUIRE(ctx, "rb_ri_dp_j: ctx is NULL");
    RB_REQUIRE(nb >= 0 && nx >= 0, "rb_ri_dp_j: negative dimension");
    RB_REQUIRE((nx == 0 || d) && (nb == 0 || j), "rb_ri_dp_j: NULL output");
    RB_CUDA(cudaSetDevice(ctx->device));
    const i64 slab = (i64)nb * nb;
    const i64 G = ctx->num_sms;
    bool fused = slab >= 4 && nx >= 1 && G <= 256 && (slab & 3) == 0 && ((((uintptr_t)ri3ao) | ((uintptr_t)dm) | ((uintptr_t)j)) & 15) == 0;
    // REST_B200_DPJ_FUSED: 0 = never, 1 = wherever the kernel can run; unset = where it was measured to win (below)
    int mode = -1;
    if (const char *e = getenv("REST_B200_DPJ_FUSED")) { if (*e) mode = atoi(e) != 0 ? 1 : 0; }
    if (mode == 0) fused = false;
    i64 L = 0, S = 0, R = 0;
    int kk = 0;
    if (fused) {
        L = (rb_cdiv(slab, G) + 3) & ~(i64)3;
        kk = (int)rb_cdiv(L, 2 * DPJ_THREADS);
        S = rb_cdiv((i64)32 * 1024, L * 8); // a block should be worth ~0.7 us of the SM's share of HBM bandwidth
        if (const char *e = getenv("REST_B200_DPJ_S")) { const i64 v = atoll(e); if (v >= 1 && v <= DPJ_MAX_S) S = v; }
        S = S >= 8 ? 8 : (S >= 4 ? 4 : (S >= 2 ? 2 : 1));
        while (S > 1 && DPJ_RING_BYTES / (S * L * 8) < 4) S >>= 1;
        R = DPJ_RING_BYTES / (S * L * 8);
        if (R > DPJ_MAX_RING) R = DPJ_MAX_RING;
        if (kk > 8 || R < 4 || G > 160) fused = false;
        // Measured (profiles/r02_dpj_fused.md): 1.42x / 1.70x / 1.29x / 1.60x faster than the two passes at nb = 600 / 700 / 800 / 900, but slower when
        // a CTA's run of a slab is short (nb = 264: 3.8 KB per run, many tiny blocks: the exchange latency of every block shows).
        // ... and nb ~ 1000 (54 KB runs: a ring of 4 one-slab blocks) measured 0.99x: no gain, so the two passes stay there as well.
        if (mode < 0 && (L * 8 < 16 * 1024 || L * 8 > 48 * 1024 || (i64)nx * slab * 8 < ((i64)256 << 20))) fused = false;
    }
    if (!fused) {
        RB_TRY(rb_ri_dp(ctx, ri3ao, dm, d, nb, nx));
        return rb_ri_j(ctx, ri3ao, d, j, nb, nx);
    }
    DpjParams p;
    p.a = ri3ao; p.slab = slab; p.nx = nx; p.dm = dm; p.d = d; p.j = j;
    // LAG: blocks a slab waits in the ring for its d_P.  The ring also has to keep ~2 blocks of loads in flight, so LAG = R - 2 is the
    // most slack the shared memory of an SM allows (R = 5 at nb = 600: LAG 3 -> 1.00 ms, 2 -> 1.13 ms).
    p.L = L; p.S = (int)S; p.R = (int)R; p.LAG = (int)(R - 2 < 3 ? R - 2 : 3);
    if (const char *e = getenv("REST_B200_DPJ_LAG")) { const int v = atoi(e); if (v >= 1 && v < p.R) p.LAG = v; }
    p.nblocks = rb_cdiv((i64)nx, S);
    p.debug = 0;
    if (const char *e = getenv("REST_B200_DPJ_DEBUG")) p.debug = atoi(e);
    const i64 partial_bytes = p.nblocks * S * G * 8;
    void *ws;
    RB_TRY(rb_ws_reserve(ctx, 1, partial_bytes, &ws));
    p.partial = (double *)ws;
    RB_CUDA(cudaMemsetAsync(p.partial, 0xff, (size_t)partial_bytes, ctx->stream)); // sentinel = "not published yet"
    const size_t smem = (size_t)(R * S * L * 8) + (size_t)(2 * DPJ_MAX_S * DPJ_WARPS + 2 * DPJ_MAX_S) * 8 + 64;
    int st;
    switch (S) {
    case 1: st = dpj_dispatch_k<1>(ctx, p, smem, kk); break;
    case 2: st = dpj_dispatch_k<2>(ctx, p, smem, kk); break;
    case 4: st = dpj_dispatch_k<4>(ctx, p, smem, kk); break;
    default: st = dpj_dispatch_k<8>(ctx, p, smem, kk); break;
    }
    if (st == DPJ_NOT_RESIDENT) {
        RB_TRY(rb_ri_dp(ctx, ri3ao, dm, d, nb, nx));
        return rb_ri_j(ctx, ri3ao, d, j, nb, nx);
    }
    return st;
}
