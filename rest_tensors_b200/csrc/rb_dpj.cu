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
// Shapes the ring cannot hold (nb > ~1100), odd nb^2 alignment or tiny problems fall back to the two GEMV kernels.
#include "rb_common.cuh"

namespace {

constexpr int DPJ_THREADS = 256;
constexpr int DPJ_WARPS = DPJ_THREADS / 32;
constexpr int DPJ_MAX_S = 32;
constexpr int DPJ_MAX_RING = 10;
constexpr i64 DPJ_RING_BYTES = 200 * 1024;

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

constexpr long long DPJ_SENTINEL = -1LL; // all ones: a NaN payload that FMA / add never produce

__device__ __forceinline__ uint32_t dpj_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// 16 bytes global -> shared without passing through registers (LDGSTS), L2 only; completion is tracked per thread in commit groups
__device__ __forceinline__ void dpj_cp16(uint32_t dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
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

// K = double2 per thread that cover the CTA's run (L <= K * 512)
template <int K>
__global__ void __launch_bounds__(DPJ_THREADS, 1) rb_ri_dp_j_kernel(const DpjParams p)
{
    extern __shared__ __align__(128) unsigned char dpj_smem[];
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const i64 G = gridDim.x, cta = blockIdx.x;
    const i64 e0 = cta * p.L;
    const int len = (int)(p.slab - e0 < p.L ? (p.slab - e0 > 0 ? p.slab - e0 : 0) : p.L); // multiple of 4 (slab and L are)
    const bool has = len > 0;
    double *ring = reinterpret_cast<double *>(dpj_smem);
    const i64 ring_elems = (i64)p.R * p.S * p.L;
    double *red = ring + ring_elems;                                                    // [S][DPJ_WARPS]
    double *dsm = red + DPJ_MAX_S * DPJ_WARPS;                                          // [S]

    // Every thread copies exactly the 16-byte units it reads back itself (unit t + k * 256 of each run), so the ring needs no barrier of
    // its own: a thread waits for its own commit groups.  One group per block, committed by every thread whether or not it copied.
    auto issue = [&](int b, int slot) {
        const int s_n = (int)(p.nx - (i64)b * p.S < p.S ? p.nx - (i64)b * p.S : p.S);
        for (int s = 0; s < s_n; ++s) {
            const double *src = p.a + ((i64)b * p.S + s) * p.slab + e0;
            const uint32_t dst = dpj_smem_u32(ring + (size_t)(slot * p.S + s) * p.L);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int q = 2 * (t + k * DPJ_THREADS);
                if (q < len) dpj_cp16(dst + (uint32_t)q * 8u, src + q);
            }
        }
        dpj_commit();
    };
    for (int b = 0; b < p.R; ++b) { // prologue: R groups (empty ones past the end keep the group arithmetic uniform)
        if (b < (int)p.nblocks) issue(b, b);
        else dpj_commit();
    }

    double2 dreg[K], jreg[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int q = 2 * (t + k * DPJ_THREADS);
        dreg[k] = (q < len) ? *reinterpret_cast<const double2 *>(p.dm + e0 + q) : make_double2(0.0, 0.0);
        jreg[k] = make_double2(0.0, 0.0);
    }

    constexpr int GV = 8; // partials per lane held in flight: up to 256 CTAs
    // loop state kept incrementally (no 64-bit divisions in the loop): ring slot / barrier parity of stage 1, ring slot of stage 2,
    // slabs left at the start of the block of each stage
    const int S = p.S, R = p.R, nblocks = (int)p.nblocks, LAG = p.LAG, Gi = (int)G;
    int slot1 = 0, slot2 = 0, left1 = (int)p.nx, left2 = (int)p.nx;
    double gv[GV]; // this warp's share of the partials of the block stage 2 handles NEXT iteration (in flight across one iteration)
#pragma unroll
    for (int i = 0; i < GV; ++i) gv[i] = 0.0;
    for (int it = 0; it < nblocks + LAG; ++it) {
        const int bb = it - LAG;
        const int s_n2 = bb >= 0 ? (left2 < S ? left2 : S) : 0;
        // ---- stage 2, first half: d_P of block bb from the partials.  Warp w gathers slabs w, w + 8, ...; the loads of the first one
        //      were issued an iteration ago, so the L2 round trip is normally over (a sentinel = not published yet: poll).
        for (int s = warp; s < s_n2; s += DPJ_WARPS) {
            if (p.debug == 1) { if (lane == 0) { dsm[s] = 1.0; if (cta == 0) p.d[bb * S + s] = 1.0; } continue; }
            const double *row = p.partial + (size_t)(bb * S + s) * Gi;
            if (s != warp) {
#pragma unroll
                for (int i = 0; i < GV; ++i) gv[i] = (lane + 32 * i < Gi) ? dpj_ld_gpu(row + lane + 32 * i) : 0.0;
            }
            const long long t0 = clock64();
            for (;;) {
                bool missing = false;
#pragma unroll
                for (int i = 0; i < GV; ++i)
                    if (__double_as_longlong(gv[i]) == DPJ_SENTINEL) { gv[i] = dpj_ld_gpu(row + lane + 32 * i); missing = true; }
                if (!__any_sync(0xffffffffu, missing)) break;
                if (clock64() - t0 > 8000000000LL) __trap(); // a lost CTA must not hang the device: fail loudly instead
            }
            double v = 0.0;
#pragma unroll
            for (int i = 0; i < GV; ++i) v += gv[i]; // lane: CTAs lane, lane + 32, ... in order; then the fixed butterfly
            v = dpj_warp_sum(v);
            if (lane == 0) {
                dsm[s] = v;
                if (cta == 0) p.d[bb * S + s] = v;
            }
        }
        // ... and ask L2 for the partials of block bb + 1 now; they are looked at in the next iteration
        if (p.debug != 1 && bb + 1 >= 0 && bb + 1 < nblocks && warp < (left2 - s_n2 < S ? left2 - s_n2 : S)) {
            const double *row = p.partial + (size_t)((bb + 1) * S + warp) * Gi;
#pragma unroll
            for (int i = 0; i < GV; ++i) gv[i] = (lane + 32 * i < Gi) ? dpj_ld_gpu(row + lane + 32 * i) : 0.0;
        }
        int slot_a = 0, s_n = 0;
        if (it < nblocks) { // ---- stage 1: partial dots of block `it`
            slot_a = slot1;
            s_n = left1 < S ? left1 : S;
            // groups committed so far: R + max(0, it - LAG); block `it` is group `it`
            dpj_wait_pending(it >= LAG ? R - LAG - 1 : R - it - 1);
            left1 -= S;
            if (++slot1 == R) slot1 = 0;
            for (int s = 0; s < s_n; ++s) {
                const double *run = ring + (size_t)(slot_a * S + s) * p.L;
                double acc = 0.0;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const int q = 2 * (t + k * DPJ_THREADS);
                    if (q < len) {
                        const double2 v = *reinterpret_cast<const double2 *>(run + q);
                        acc += dreg[k].x * v.x;
                        acc += dreg[k].y * v.y;
                    }
                }
                acc = dpj_warp_sum(acc);
                if (lane == 0) red[s * DPJ_WARPS + warp] = acc;
            }
        }
        __syncthreads(); // red (stage 1) and dsm (stage 2) are complete
        if (t < s_n) {
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < DPJ_WARPS; ++w) v += red[t * DPJ_WARPS + w];
            dpj_st_gpu(p.partial + (size_t)(it * S + t) * Gi + cta, v);
        }
        if (bb >= 0) { // ---- stage 2, second half: J += d_P * A from the copy still in the ring
            const int slot = slot2;
            left2 -= S;
            if (++slot2 == R) slot2 = 0;
            if (has) {
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const int q = 2 * (t + k * DPJ_THREADS);
                    if (q < len) {
                        for (int s = 0; s < s_n2; ++s) {
                            const double2 v = *reinterpret_cast<const double2 *>(ring + (size_t)(slot * S + s) * p.L + q);
                            const double ds = dsm[s];
                            jreg[k].x += ds * v.x;
                            jreg[k].y += ds * v.y;
                        }
                    }
                }
            }
            // the slot is private per thread (each thread reads only what it copied), so it can be refilled right away
            if (bb + R < nblocks) issue(bb + R, slot);
            else dpj_commit();
            __syncthreads(); // every thread is done with dsm and with red
        } else {
            __syncthreads(); // red is reused by the next iteration's stage 1
        }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int q = 2 * (t + k * DPJ_THREADS);
        if (q < len) *reinterpret_cast<double2 *>(p.j + e0 + q) = jreg[k];
    }
}

template <int K>
int dpj_launch(rb_ctx *ctx, const DpjParams &p, size_t smem)
{
    static int configured_for = -1; // per process and device: the attribute belongs to the function on that device
    if (configured_for != ctx->device) {
        RB_CUDA(cudaFuncSetAttribute(rb_ri_dp_j_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(DPJ_RING_BYTES + 8192)));
        configured_for = ctx->device;
    }
    void *args[] = {(void *)&p};
    RB_CUDA(cudaLaunchCooperativeKernel((const void *)rb_ri_dp_j_kernel<K>, dim3((unsigned)ctx->num_sms), dim3(DPJ_THREADS), args, smem,
                                        ctx->stream));
    RB_LAUNCHED(ctx);
    return RB_OK;
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
    // Opt-in: measured at config C (tools/prof_dpj.py, profiles/r02_dpj_fused.md) the single pass takes 2.0 ms against 1.46 ms for the two
    // GEMV passes -- DRAM traffic is halved as designed (4.90 GB), but with one 256-thread CTA per SM the loop is bound by its own
    // instruction latency (ncu: no memory stalls; 4 600 cycles per 2-slab block), not by HBM.
    {
        const char *e = getenv("REST_B200_DPJ_FUSED");
        fused = fused && e && atoi(e) != 0;
    }
    i64 L = 0, S = 0, R = 0;
    int kk = 0;
    if (fused) {
        L = (rb_cdiv(slab, G) + 3) & ~(i64)3;
        kk = (int)rb_cdiv(L, 2 * DPJ_THREADS);
        S = rb_cdiv((i64)20 * 1024, L * 8); // a block should be worth ~0.4 us of the SM's share of HBM bandwidth
        if (S > DPJ_MAX_S) S = DPJ_MAX_S;
        if (const char *e = getenv("REST_B200_DPJ_S")) { const i64 v = atoll(e); if (v >= 1 && v <= DPJ_MAX_S) S = v; }
        if (S > nx) S = nx;
        if (S < 1) S = 1;
        R = DPJ_RING_BYTES / (S * L * 8);
        if (R > DPJ_MAX_RING) R = DPJ_MAX_RING;
        if (kk > 16 || R < 4) fused = false;
    }
    if (!fused) {
        RB_TRY(rb_ri_dp(ctx, ri3ao, dm, d, nb, nx));
        return rb_ri_j(ctx, ri3ao, d, j, nb, nx);
    }
    DpjParams p;
    p.a = ri3ao; p.slab = slab; p.nx = nx; p.dm = dm; p.d = d; p.j = j;
    p.L = L; p.S = (int)S; p.R = (int)R; p.LAG = (int)(R / 2);
    if (const char *e = getenv("REST_B200_DPJ_LAG")) { const int v = atoi(e); if (v >= 1 && v < p.R) p.LAG = v; }
    p.nblocks = rb_cdiv((i64)nx, S);
    p.debug = 0;
    if (const char *e = getenv("REST_B200_DPJ_DEBUG")) p.debug = atoi(e);
    const i64 partial_bytes = p.nblocks * S * G * 8;
    void *ws;
    RB_TRY(rb_ws_reserve(ctx, 1, partial_bytes, &ws));
    p.partial = (double *)ws;
    RB_CUDA(cudaMemsetAsync(p.partial, 0xff, (size_t)partial_bytes, ctx->stream)); // sentinel = "not published yet"
    const size_t smem = (size_t)(R * S * L * 8) + 128 + (size_t)(DPJ_MAX_S * DPJ_WARPS + DPJ_MAX_S) * 8;
    switch (kk) {
    case 1: return dpj_launch<1>(ctx, p, smem);
    case 2: return dpj_launch<2>(ctx, p, smem);
    case 3: return dpj_launch<3>(ctx, p, smem);
    case 4: return dpj_launch<4>(ctx, p, smem);
    case 5: return dpj_launch<5>(ctx, p, smem);
    case 6: return dpj_launch<6>(ctx, p, smem);
    case 7: case 8: return dpj_launch<8>(ctx, p, smem);
    case 9: case 10: return dpj_launch<10>(ctx, p, smem);
    case 11: case 12: return dpj_launch<12>(ctx, p, smem);
    default: return dpj_launch<16>(ctx, p, smem);
    }
}
