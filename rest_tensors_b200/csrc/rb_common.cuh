// rb_common.cuh -- context, error plumbing and small device helpers shared by all translation units.
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <cstdlib>
#include <mutex>
#include <vector>
#include "../../include/rest_b200.h"

typedef int64_t i64;

// ---- error plumbing ---------------------------------------------------------------------------------
void rb_set_error(const char *fmt, ...);

#define RB_CUDA(call)                                                                                   \
    do {                                                                                                \
        cudaError_t _e = (call);                                                                        \
        if (_e != cudaSuccess) {                                                                        \
            rb_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e));         \
            (void)cudaGetLastError(); /* reported here: must not resurface in a later launch check */   \
            return RB_ERR_CUDA;                                                                         \
        }                                                                                               \
    } while (0)

#define RB_REQUIRE(cond, ...)                                                                           \
    do {                                                                                                \
        if (!(cond)) {                                                                                  \
            rb_set_error(__VA_ARGS__);                                                                  \
            return RB_ERR_INVALID;                                                                      \
        }                                                                                               \
    } while (0)

#define RB_TRY(expr)                                                                                    \
    do {                                                                                                \
        int _s = (expr);                                                                                \
        if (_s != RB_OK) return _s;                                                                     \
    } while (0)

// ---- context ----------------------------------------------------------------------------------------
typedef CUresult (*rb_encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                       const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                       CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                       CUtensorMapFloatOOBfill);

struct rb_ctx {
    int device = 0;
    int num_sms = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr; // the stream calls run on (own_stream or the caller's)
    void *ws[4] = {nullptr, nullptr, nullptr, nullptr}; // grow-only workspaces (slot 0: RI ops, 1: GEMM split-K, 2: host staging, 3: probes)
    i64 ws_bytes[4] = {0, 0, 0, 0};
    i64 ws_budget = 0;             // cached workspace budget (bytes), 0 = not yet queried
    i64 launches = 0;
    i64 tma_layout_launches = 0;   // launches of the bulk-tensor layout kernels (rb_layout_tma.cu)
    int gemm_path = 0;
    int layout_path = 0;           // 0: 256-bit / plain-load layout kernels, 1: bulk-tensor (TMA) copy / transpose where eligible
    rb_encode_tiled_fn encode_tiled = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaStream_t aux_stream = nullptr;   // copy stream of the peer pipeline (rb_ri_mo_pq_peers), created on first use
    cudaEvent_t aux_ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr}; // [0] fork, [1..2] pulled, [3..4] buffer free
    void *eig_cache = nullptr;           // instantiated Jacobi sweep graphs (rb_eig.cu), freed by rb_eig_cache_free
    unsigned long long *sched = nullptr; // GEMM tile-scheduler slots (64 x 2 words on the device), zero between launches
    unsigned sched_next = 0;             // slot of the next GEMM launch (round-robin)
    void *comm = nullptr;                // NCCL communicator (rb_comm.cu), NULL = a world of one
    int comm_rank = 0, comm_world = 1;
    // CUDA-graph capture of a call sequence (rb_graph_begin / rb_graph_end): while `capturing`, calls that need the host in the loop
    // (workspace growth, eigen-solvers, collectives) are refused; while recordings are alive, a workspace block that has to grow is
    // parked instead of freed (the recordings hold its address).
    int capturing = 0;
    int live_graphs = 0;
    i64 capture_launches0 = 0;
    std::vector<void *> ws_parked;
};

// For entry points that synchronise with the host or allocate: they cannot be recorded into a CUDA graph.
#define RB_NO_CAPTURE(ctx, name)                                                                        \
    do {                                                                                                \
        if ((ctx)->capturing) {                                                                         \
            rb_set_error("%s cannot be recorded into a graph (it needs the host inside the call): "     \
                         "issue it outside rb_graph_begin / rb_graph_end", name);                       \
            return RB_ERR_UNSUPPORTED;                                                                  \
        }                                                                                               \
    } while (0)

// Grow-only device workspace (synchronises the stream before freeing the old block).
int rb_ws_reserve(rb_ctx *ctx, int slot, i64 bytes, void **out);

// After every kernel launch: count it and surface launch-configuration errors.
#define RB_LAUNCHED(ctx)                                                                                \
    do {                                                                                                \
        (ctx)->launches++;                                                                              \
        cudaError_t _e = cudaGetLastError();                                                            \
        if (_e != cudaSuccess) {                                                                        \
            rb_set_error("%s:%d: kernel launch failed -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return RB_ERR_CUDA;                                                                         \
        }                                                                                               \
    } while (0)

static inline i64 rb_cdiv(i64 a, i64 b) { return (a + b - 1) / b; }
static inline bool rb_is_n(char c) { return c == 'N' || c == 'n'; }
static inline bool rb_is_t(char c) { return c == 'T' || c == 't'; }
static inline bool rb_is_u(char c) { return c == 'U' || c == 'u'; }
static inline bool rb_is_l(char c) { return c == 'L' || c == 'l'; }

// ---- 256-bit global accesses (sm_100: LDG.E.ENL2.256 / STG.E.ENL2.256) ------------------------------------------------
// Measured on this pool's B200 (tools/micro/copy_bench.cu, profiles/r02_copy_microbench.md): a 1 read : 1 write SM kernel
// moves 6.6-6.7 TB/s with 32-byte accesses, 8 of them in flight per thread and >= 16 CTAs per SM in the grid, against 6.3 TB/s
// with 16-byte accesses and 6.5 TB/s for the driver's device-to-device memcpy.  Addresses must be 32-byte aligned.
struct alignas(32) rb_d4 { double x, y, z, w; };
#ifdef __CUDACC__
__device__ __forceinline__ rb_d4 rb_ld256(const double *p)
{
    rb_d4 v;
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void rb_st256(double *p, const rb_d4 &v)
{
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
}
#endif
static inline bool rb_aligned32(const void *p) { return (((uintptr_t)p) & 31) == 0; }

// ---- internal cross-TU entry points -------------------------------------------------------------------
// Generic strided 3-D copy: dst[d0 + i*di + j*dj + k*dk] = src[s0 + i*si + j*sj + k*sk]
int rb_copy3d(rb_ctx *ctx, const double *src, i64 s0, i64 si, i64 sj, i64 sk, double *dst, i64 d0, i64 di, i64 dj,
              i64 dk, i64 ni, i64 nj, i64 nk);
// Batched tiled 2-D transpose: out[c + r*ors + b*obs] = in[r + c*ics + b*ibs]
int rb_transpose_batched(rb_ctx *ctx, const double *in, i64 ics, i64 ibs, double *out, i64 ors, i64 obs, i64 nr,
                         i64 nc, i64 nbatch);
// TMA forms (rb_layout_tma.cu); RB_TMA_NOT_ELIGIBLE (-1): operands TMA cannot describe, the caller runs the plain kernel
#define RB_TMA_NOT_ELIGIBLE (-1)
int rb_tma_copy3d(rb_ctx *ctx, const double *s, i64 sj, i64 sk, double *d, i64 dj, i64 dk, i64 ni, i64 nj, i64 nk);
int rb_tma_transpose(rb_ctx *ctx, const double *in, i64 ics, i64 ibs, double *out, i64 ors, i64 obs, i64 nr, i64 nc, i64 nbatch);
// GEMM core (rb_gemm.cu). tri: 0 = full, 1 = only tiles/elements with row<=col (upper), 2 = row>=col (lower)
int rb_gemm_core(rb_ctx *ctx, bool ta, bool tb, i64 m, i64 n, i64 k, double alpha, const double *a, i64 lda,
                 i64 stride_a, const double *b, i64 ldb, i64 stride_b, double beta, double *c, i64 ldc,
                 i64 stride_c, i64 batch, int tri);
// Symmetrise: copy the `uplo` triangle of the n x n matrix c onto the other triangle.
int rb_symmetrize(rb_ctx *ctx, double *c, i64 n, i64 ldc, bool from_upper);
// y = beta-scaled / zero helper
int rb_scale_or_zero(rb_ctx *ctx, double *y, i64 n, i64 inc, double beta);

// rb_ri.cu: upper triangle of k (+)= sum_P (A_P ct)(A_P ct)^T over nx slabs (beta 0 overwrite / 1 accumulate)
int rb_ri_k_upper(rb_ctx *ctx, const double *ri3ao, const double *ct, i64 no, double *k, i64 nb, i64 nx, double beta);

// rb_eri.cu: ERIFold4 chunk scatter without bounds checks (dst may be a virtual origin of a window)
int rb_erifold4_scatter(rb_ctx *ctx, double *dst, i64 ld, const double *buf, i64 i0, i64 li, i64 j0, i64 lj, i64 k0, i64 lk, i64 l0,
                        i64 ll, int mode);

// rb_eig.cu: drop the cached sweep graphs of a context (called by rb_ctx_destroy)
void rb_eig_cache_free(rb_ctx *ctx);

// Default (process-wide) context for the host-pointer entry points.
rb_ctx *rb_default_ctx(void);
std::mutex &rb_default_mutex(void);
