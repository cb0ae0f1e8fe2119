// rb_compat.cu -- HOST-pointer entry points.
//   (1) the seven Fortran-ABI symbols the reference's Rust FFI binds (src/external_libs/ffi_restmatr.rs:4-62),
//   (2) rb_host_* mirrors of the reference's BLAS / layout calls (src/matrix/matrix_blas_lapack.rs, matrixupper.rs,
//       matrixfull.rs, ri.rs).
// Every call stages its operands into HBM, runs the same CUDA kernels as the device API and copies the result back
// before returning; nothing is retained.  There is no CPU arithmetic here: the only host work is moving bytes
// (cudaMemcpy, and memcpy into / out of pinned bounce blocks when the caller's buffers are pageable).
// ri_ao2mo_f_ / rb_host_ri_ao2mo stream P-chunks through a 3-stream pipeline (H2D | DMMA GEMMs | D2H) so that PCIe
// transfers overlap the contraction.  Pinned caller buffers (rb_host_alloc_pinned) go straight to the DMA engines
// (config C: ~120-130 ms per pass); pageable ones (a Rust Vec<f64>) are bounced by host threads (~250 ms instead of
// the ~720 ms the driver's synchronous staged copies take).
#include "rb_common.cuh"
#include <vector>
#include <string>
#ifndef RB_DEFAULT_HEAD
#define RB_DEFAULT_HEAD ""
#endif
#include <chrono>
#include <atomic>
#include <thread>
#include <cstring>
#include <sched.h>

static double now_ms()
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

namespace {

// Device staging blocks of the host-pointer entry points are cached between calls (cudaMalloc/cudaFree of GB-sized
// buffers cost tens of ms per call otherwise).  Calls are serialised by the default-context mutex, so a plain list is
// enough; blocks beyond the cap are released when a call ends.
struct StageBlock { void *p; i64 bytes; bool busy; };
static std::vector<StageBlock> g_stage;
static const i64 STAGE_CACHE_CAP = (i64)24 << 30;

// ---- pageable caller buffers -----------------------------------------------------------------------------------
// A Rust Vec<f64> (what the reference's RIFull / MatrixFull own) is pageable: cudaMemcpyAsync on it is staged by the
// driver and synchronous, which serialises the 3-stream pipeline (measured: 740 ms instead of 132 ms per config-C
// pass).  For such buffers the streaming pass bounces every chunk through cached PINNED blocks: host threads copy
// chunk i+1 in / scatter chunk i-1 out (plain memcpy -- data movement, no arithmetic) while the GPU works on chunk i.
struct PinBlock { void *p; i64 bytes; bool busy; };
static std::vector<PinBlock> g_pin;
static const i64 PIN_CACHE_CAP = (i64)8 << 30;

static bool host_ptr_is_pinned(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}

static void *pin_take(i64 bytes)
{
    int best = -1;
    for (size_t i = 0; i < g_pin.size(); ++i)
        if (!g_pin[i].busy && g_pin[i].bytes >= bytes && (best < 0 || g_pin[i].bytes < g_pin[best].bytes)) best = (int)i;
    if (best >= 0) { g_pin[best].busy = true; return g_pin[best].p; }
    i64 total = 0;
    for (auto &b : g_pin) total += b.bytes;
    for (size_t i = 0; i < g_pin.size() && total + bytes > PIN_CACHE_CAP;) { // make room: drop idle blocks
        if (!g_pin[i].busy) { cudaFreeHost(g_pin[i].p); total -= g_pin[i].bytes; g_pin.erase(g_pin.begin() + i); } else ++i;
    }
    void *p = nullptr;
    if (cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    g_pin.push_back({p, bytes, true});
    return p;
}

static void pin_release_all()
{
    for (auto &b : g_pin) b.busy = false;
}

static int host_threads()
{
    static int n = 0;
    if (n) return n;
    int have = 1;
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof set, &set) == 0) have = CPU_COUNT(&set);
    int t = have > 16 ? 16 : have; // a transient burst of memcpy threads; measured 8 -> 16 threads: 279 -> 248 ms per config-C pass
    if (const char *e = getenv("REST_B200_HOST_THREADS")) t = atoi(e);
    n = t < 1 ? 1 : (t > 64 ? 64 : t);
    return n;
}

// run fn(task) for task in [0, ntasks) on host_threads() threads (the caller is one of them)
template <typename F>
static void parallel_tasks(i64 ntasks, F fn)
{
    const int nt = (int)(ntasks < host_threads() ? ntasks : host_threads());
    std::atomic<i64> next(0);
    auto work = [&]() { for (i64 t = next.fetch_add(1); t < ntasks; t = next.fetch_add(1)) fn(t); };
    std::vector<std::thread> th;
    for (int i = 1; i < nt; ++i) th.emplace_back(work);
    work();
    for (auto &x : th) x.join();
}

// dst[r*dpitch .. +width) = src[r*spitch .. +width) for r in [0, rows): split into ~4 MB tasks
static void host_copy_2d(double *dst, i64 dpitch, const double *src, i64 spitch, i64 width, i64 rows)
{
    if (width <= 0 || rows <= 0) return;
    if (dpitch == width && spitch == width) { // contiguous: split by bytes
        const i64 total = width * rows, piece = (i64)1 << 19; // 4 MB of doubles
        parallel_tasks((total + piece - 1) / piece, [&](i64 t) {
            const i64 o = t * piece, n = (total - o < piece) ? total - o : piece;
            memcpy(dst + o, src + o, (size_t)n * 8);
        });
        return;
    }
    i64 rpt = ((i64)1 << 19) / width;
    if (rpt < 1) rpt = 1;
    parallel_tasks((rows + rpt - 1) / rpt, [&](i64 t) {
        const i64 r0 = t * rpt, r1 = (r0 + rpt < rows) ? r0 + rpt : rows;
        for (i64 r = r0; r < r1; ++r) memcpy(dst + r * dpitch, src + r * spitch, (size_t)width * 8);
    });
}

struct HostOp {
    rb_ctx *ctx;
    std::unique_lock<std::mutex> lock;
    HostOp() : ctx(nullptr), lock(rb_default_mutex()) { ctx = rb_default_ctx(); }
    ~HostOp()
    {
        if (ctx) cudaStreamSynchronize(ctx->stream);
        for (int i = 0; i < 2; ++i) if (piece_ev[i]) cudaEventDestroy(piece_ev[i]);
        i64 total = 0;
        pin_release_all();
        for (auto &b : g_stage) { b.busy = false; total += b.bytes; }
        while (total > STAGE_CACHE_CAP && !g_stage.empty()) { // drop the largest blocks first
            size_t big = 0;
            for (size_t i = 1; i < g_stage.size(); ++i) if (g_stage[i].bytes > g_stage[big].bytes) big = i;
            cudaFree(g_stage[big].p);
            total -= g_stage[big].bytes;
            g_stage.erase(g_stage.begin() + big);
        }
    }
    int alloc(i64 elems, double **out)
    {
        *out = nullptr;
        if (elems <= 0) elems = 1;
        const i64 need = elems * 8;
        int best = -1;
        for (size_t i = 0; i < g_stage.size(); ++i) // best fit among free cached blocks (at most 2x oversize)
            if (!g_stage[i].busy && g_stage[i].bytes >= need && g_stage[i].bytes <= 2 * need + 4096 &&
                (best < 0 || g_stage[i].bytes < g_stage[best].bytes)) best = (int)i;
        if (best >= 0) { g_stage[best].busy = true; *out = (double *)g_stage[best].p; return RB_OK; }
        void *p = nullptr;
        cudaError_t e = cudaMalloc(&p, (size_t)need);
        if (e != cudaSuccess) { // release the idle cache and retry once
            cudaGetLastError();
            for (size_t i = 0; i < g_stage.size();) {
                if (!g_stage[i].busy) { cudaFree(g_stage[i].p); g_stage.erase(g_stage.begin() + i); } else ++i;
            }
            e = cudaMalloc(&p, (size_t)need);
        }
        if (e != cudaSuccess) {
            cudaGetLastError();
            rb_set_error("host wrapper: cudaMalloc(%lld doubles) failed: %s", (long long)elems, cudaGetErrorString(e));
            return RB_ERR_NOMEM;
        }
        g_stage.push_back({p, need, true});
        *out = (double *)p;
        return RB_OK;
    }
    // upload a host matrix block [rows x cols] with leading dimension ld into a dense device buffer
    // ---- large pageable operands: pipelined bounce through two pinned pieces (memcpy threads | DMA) ----
    static constexpr i64 PIECE = (i64)8 << 20;     // doubles per pinned piece (64 MB)
    static constexpr i64 BOUNCE_MIN = (i64)2 << 20; // doubles (16 MB): below this the driver's staged copy is fine
    double *piece[2] = {nullptr, nullptr};
    cudaEvent_t piece_ev[2] = {nullptr, nullptr};
    bool want_bounce(const void *host, i64 n)
    {
        if (n < BOUNCE_MIN) return false;
        if (const char *e = getenv("REST_B200_BOUNCE")) if (atoi(e) == 0) return false;
        if (host_ptr_is_pinned(host)) return false;
        for (int i = 0; i < 2; ++i) {
            if (!piece[i]) piece[i] = (double *)pin_take(PIECE * 8);
            if (!piece[i]) return false;
            if (!piece_ev[i] && cudaEventCreateWithFlags(&piece_ev[i], cudaEventDisableTiming) != cudaSuccess) return false;
        }
        return true;
    }
    // host [rows x cols] with pitch ld  ->  dense device [rows x cols]
    int up2d(double *dst, const double *src, i64 rows, i64 cols, i64 ld)
    {
        if (rows <= 0 || cols <= 0) return RB_OK;
        if (!want_bounce(src, rows * cols)) {
            RB_CUDA(cudaMemcpy2DAsync(dst, (size_t)rows * 8, src, (size_t)ld * 8, (size_t)rows * 8, (size_t)cols,
                                      cudaMemcpyHostToDevice, ctx->stream));
            return RB_OK;
        }
        i64 cpp = PIECE / rows; // whole columns per piece
        if (cpp < 1) { // a single column longer than a piece: split it
            for (i64 c = 0; c < cols; ++c) RB_TRY(up_dense(dst + c * rows, src + c * ld, rows));
            return RB_OK;
        }
        int k = 0;
        for (i64 c0 = 0; c0 < cols; c0 += cpp, ++k) {
            const int s = k & 1;
            const i64 cn = cols - c0 < cpp ? cols - c0 : cpp;
            if (k >= 2) RB_CUDA(cudaEventSynchronize(piece_ev[s]));
            host_copy_2d(piece[s], rows, src + c0 * ld, ld, rows, cn);
            RB_CUDA(cudaMemcpyAsync(dst + c0 * rows, piece[s], (size_t)(rows * cn) * 8, cudaMemcpyHostToDevice, ctx->stream));
            RB_CUDA(cudaEventRecord(piece_ev[s], ctx->stream));
        }
        RB_CUDA(cudaEventSynchronize(piece_ev[0]));
        if (k > 1) RB_CUDA(cudaEventSynchronize(piece_ev[1])); // the pieces may be reused by the next transfer
        return RB_OK;
    }
    int up_dense(double *dst, const double *src, i64 n)
    {
        int k = 0;
        for (i64 o = 0; o < n; o += PIECE, ++k) {
            const int s = k & 1;
            const i64 len = n - o < PIECE ? n - o : PIECE;
            if (k >= 2) RB_CUDA(cudaEventSynchronize(piece_ev[s]));
            host_copy_2d(piece[s], len, src + o, len, len, 1);
            RB_CUDA(cudaMemcpyAsync(dst + o, piece[s], (size_t)len * 8, cudaMemcpyHostToDevice, ctx->stream));
            RB_CUDA(cudaEventRecord(piece_ev[s], ctx->stream));
        }
        RB_CUDA(cudaEventSynchronize(piece_ev[0]));
        if (k > 1) RB_CUDA(cudaEventSynchronize(piece_ev[1]));
        return RB_OK;
    }
    // dense device [rows x cols]  ->  host with pitch ld.  The bounced form returns with the data in place.
    int down2d(double *dst, i64 ld, const double *src, i64 rows, i64 cols)
    {
        if (rows <= 0 || cols <= 0) return RB_OK;
        if (!want_bounce(dst, rows * cols)) {
            RB_CUDA(cudaMemcpy2DAsync(dst, (size_t)ld * 8, src, (size_t)rows * 8, (size_t)rows * 8, (size_t)cols,
                                      cudaMemcpyDeviceToHost, ctx->stream));
            return RB_OK;
        }
        i64 cpp = PIECE / rows;
        if (cpp < 1) {
            for (i64 c = 0; c < cols; ++c) RB_TRY(down_dense(dst + c * ld, src + c * rows, rows));
            return RB_OK;
        }
        int k = 0;
        i64 prev_c0 = 0, prev_cn = 0;
        for (i64 c0 = 0; c0 < cols; c0 += cpp, ++k) {
            const int s = k & 1;
            const i64 cn = cols - c0 < cpp ? cols - c0 : cpp;
            RB_CUDA(cudaMemcpyAsync(piece[s], src + c0 * rows, (size_t)(rows * cn) * 8, cudaMemcpyDeviceToHost, ctx->stream));
            RB_CUDA(cudaEventRecord(piece_ev[s], ctx->stream));
            if (k >= 1) { // scatter the previous piece while this one is in flight
                RB_CUDA(cudaEventSynchronize(piece_ev[s ^ 1]));
                host_copy_2d(dst + prev_c0 * ld, ld, piece[s ^ 1], rows, rows, prev_cn);
            }
            prev_c0 = c0; prev_cn = cn;
        }
        RB_CUDA(cudaEventSynchronize(piece_ev[(k - 1) & 1]));
        host_copy_2d(dst + prev_c0 * ld, ld, piece[(k - 1) & 1], rows, rows, prev_cn);
        return RB_OK;
    }
    int down_dense(double *dst, const double *src, i64 n)
    {
        int k = 0;
        i64 prev_o = 0, prev_len = 0;
        for (i64 o = 0; o < n; o += PIECE, ++k) {
            const int s = k & 1;
            const i64 len = n - o < PIECE ? n - o : PIECE;
            RB_CUDA(cudaMemcpyAsync(piece[s], src + o, (size_t)len * 8, cudaMemcpyDeviceToHost, ctx->stream));
            RB_CUDA(cudaEventRecord(piece_ev[s], ctx->stream));
            if (k >= 1) {
                RB_CUDA(cudaEventSynchronize(piece_ev[s ^ 1]));
                host_copy_2d(dst + prev_o, prev_len, piece[s ^ 1], prev_len, prev_len, 1);
            }
            prev_o = o; prev_len = len;
        }
        RB_CUDA(cudaEventSynchronize(piece_ev[(k - 1) & 1]));
        host_copy_2d(dst + prev_o, prev_len, piece[(k - 1) & 1], prev_len, prev_len, 1);
        return RB_OK;
    }
    int up(double *dst, const double *src, i64 n)
    {
        if (n <= 0) return RB_OK;
        if (want_bounce(src, n)) return up_dense(dst, src, n);
        RB_CUDA(cudaMemcpyAsync(dst, src, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
        return RB_OK;
    }
    int down(double *dst, const double *src, i64 n)
    {
        if (n <= 0) return RB_OK;
        if (want_bounce(dst, n)) return down_dense(dst, src, n);
        RB_CUDA(cudaMemcpyAsync(dst, src, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
        return RB_OK;
    }
    int sync()
    {
        RB_CUDA(cudaStreamSynchronize(ctx->stream));
        return RB_OK;
    }
};

} // namespace

// Release the cached device staging blocks and pinned bounce blocks of the host-pointer entry points.
extern "C" int rb_host_trim(void)
{
    std::unique_lock<std::mutex> lock(rb_default_mutex());
    rb_ctx *ctx = rb_default_ctx();
    if (ctx) cudaStreamSynchronize(ctx->stream);
    for (auto &b : g_stage) cudaFree(b.p);
    g_stage.clear();
    for (auto &b : g_pin) cudaFreeHost(b.p);
    g_pin.clear();
    cudaGetLastError();
    return RB_OK;
}

namespace {

#define HOST_CTX(op)                                                                                     \
    HostOp op;                                                                                           \
    if (!op.ctx) return RB_ERR_CUDA

void die_if(int status, const char *who)
{
    if (status != RB_OK) {
        fprintf(stderr, "librest_b200: %s failed: %s\n", who, rb_last_error());
        abort();
    }
}

// ---- pipelined host ao2mo ---------------------------------------------------------------------------------------
struct Pipe {
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t in_done[2] = {nullptr, nullptr}, comp_done[2] = {nullptr, nullptr}, out_done[2] = {nullptr, nullptr};
    ~Pipe()
    {
        // an early return (e.g. a failed workspace allocation mid-pipeline) may leave copies in flight on the side streams
        // that read / write the stage and pinned blocks HostOp is about to hand back: drain them first
        if (s_in) cudaStreamSynchronize(s_in);
        if (s_out) cudaStreamSynchronize(s_out);
        for (int i = 0; i < 2; ++i) {
            if (in_done[i]) cudaEventDestroy(in_done[i]);
            if (comp_done[i]) cudaEventDestroy(comp_done[i]);
            if (out_done[i]) cudaEventDestroy(out_done[i]);
        }
        if (s_in) cudaStreamDestroy(s_in);
        if (s_out) cudaStreamDestroy(s_out);
    }
    int init()
    {
        RB_CUDA(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
        RB_CUDA(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            RB_CUDA(cudaEventCreateWithFlags(&in_done[i], cudaEventDisableTiming));
            RB_CUDA(cudaEventCreateWithFlags(&comp_done[i], cudaEventDisableTiming));
            RB_CUDA(cudaEventCreateWithFlags(&out_done[i], cudaEventDisableTiming));
        }
        return RB_OK;
    }
};

// One streaming pass over the host ri3ao: every P-chunk is uploaded ONCE and feeds ao2mo (if out != NULL) and the
// d_P / J / K builds (if dm / ct != NULL); J and K accumulate across chunks on the device.
// ri3mo[P, a, b] -> the a <= b part, P fastest: packed[P + pn * (b (b + 1) / 2 + a)].  For a fixed b the pairs (0..b, b) are one
// contiguous run of (b + 1) * pn doubles on both sides, so this is nl plain copies (blockIdx.y = b); 32-byte accesses when pn % 4 == 0.
__global__ void __launch_bounds__(256) rb_ri3mo_pack_pairs_kernel(const double *__restrict__ full, double *__restrict__ packed, i64 pn, i64 nl,
                                                                 int vec4)
{
    const i64 b = blockIdx.y;
    const i64 len = (b + 1) * pn;
    const double *s = full + pn * nl * b;
    double *d = packed + pn * (b * (b + 1) / 2);
    const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x, nt = (i64)gridDim.x * blockDim.x;
    if (vec4) {
        const i64 n4 = len >> 2;
        i64 i = t;
        for (; i + nt < n4; i += 2 * nt) {
            const rb_d4 v0 = rb_ld256(s + 4 * i), v1 = rb_ld256(s + 4 * (i + nt));
            rb_st256(d + 4 * i, v0); rb_st256(d + 4 * (i + nt), v1);
        }
        for (; i < n4; i += nt) rb_st256(d + 4 * i, rb_ld256(s + 4 * i));
    } else {
        for (i64 i = t; i < len; i += nt) d[i] = s[i];
    }
}

// Slabs uploaded as upper trapezoids only (symm_in): every slab of the chunk gets its strict lower triangle from its upper one.
// 32 x 32 tiles through shared memory, blockIdx.y = slab; both sides coalesced.
__global__ void __launch_bounds__(256) rb_mirror_upper_slabs_kernel(double *__restrict__ a, i64 n, i64 slab)
{
    __shared__ double tile[32][33];
    i64 p = blockIdx.x;
    i64 tj = (i64)((sqrt(8.0 * (double)p + 1.0) - 1.0) * 0.5);
    while (tj * (tj + 1) / 2 > p) --tj;
    while ((tj + 1) * (tj + 2) / 2 <= p) ++tj;
    const i64 ti = p - tj * (tj + 1) / 2; // ti <= tj: source tile rows ti, cols tj
    double *c = a + (i64)blockIdx.y * slab;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int cc = ty + r * 8;
        const i64 row = ti * 32 + tx, col = tj * 32 + cc;
        tile[cc][tx] = (row < n && col < n && row <= col) ? c[row + col * n] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int cc = ty + r * 8;
        const i64 row = tj * 32 + tx, col = ti * 32 + cc; // transposed position
        if (row < n && col < n && row > col) c[row + col * n] = tile[tx][cc];
    }
}

int host_ri_stream(const double *cl, int nl, const double *cr, int nr, const double *ri3ao, double *out, int nb_, int nx_,
                   const double *dm, const double *ct, int no, double *d_out, double *j_out, double *k_out, bool upper_out = false,
                   bool symm_in = false)
{
    RB_REQUIRE(nl >= 0 && nr >= 0 && nb_ >= 0 && nx_ >= 0 && no >= 0, "ri stream: negative dimension");
    RB_REQUIRE(!upper_out || (cl == cr && nl == nr), "ri stream: the packed (a <= b) output needs the same C on both sides");
    const i64 nb = nb_, nx = nx_;
    const bool do_mo = out != nullptr && nl > 0 && nr > 0;
    const bool do_j = dm != nullptr && (d_out != nullptr || j_out != nullptr);
    const bool do_k = ct != nullptr && k_out != nullptr && no > 0;
    if (nx == 0 || !(do_mo || do_j || do_k)) return RB_OK;
    const bool trace = getenv("REST_B200_TRACE") != nullptr;
    const double t_start = now_ms();
    HOST_CTX(op);
    rb_ctx *ctx = op.ctx;
    // chunk of P slabs staged per pipeline step: large enough for full 128-row MMA tiles along P, small enough
    // to overlap the PCIe transfers with compute
    // 256 slabs per chunk for long tensors; short ones (configs A / B: 400 / 720 slabs) get 96 so that the pipeline still has 4-8 chunks to
    // overlap (measured, tools/e2e_small_probe.py: config A 1.41 -> 1.20 ms, config B 12.0 -> 10.1 ms per pass)
    i64 pc = nx > 1024 ? 256 : 96;
    if (const char *e = getenv("REST_B200_PC")) { i64 v = atoll(e); if (v >= 8) pc = v; }
    // (Measured on this pool, profiles/r01_e2e_variants.md: letting GEMM 2 store straight into mapped pinned host memory
    //  is slower -- 223 ms vs 130 ms per pass at config C -- than staging in HBM and draining with 2-D copies.)
    // slab_out: columns of the device ri3mo chunk; ship_out: columns that travel to the host (the a <= b pairs when upper_out)
    const i64 slab_in = nb * nb, slab_out = do_mo ? (i64)nl * nr : 0;
    const i64 ship_out = (do_mo && upper_out) ? (i64)nl * (nl + 1) / 2 : slab_out;
    while (pc > 8 && pc * (slab_in + slab_out + (upper_out ? ship_out : 0)) * 8 * 2 > ((i64)6 << 30)) pc >>= 1;
    if (pc > nx) pc = nx;
    double *d_cl = nullptr, *d_cr = nullptr, *d_in[2], *d_mo[2] = {nullptr, nullptr}, *d_pk[2] = {nullptr, nullptr};
    double *d_dm = nullptr, *d_ct = nullptr, *d_d = nullptr, *d_j = nullptr, *d_k = nullptr;
    const bool same_c = (cl == cr && nl == nr);
    if (do_mo) {
        RB_TRY(op.alloc(nb * nl, &d_cl));
        if (same_c) d_cr = d_cl; else RB_TRY(op.alloc(nb * nr, &d_cr));
    }
    for (int i = 0; i < 2; ++i) {
        RB_TRY(op.alloc(pc * slab_in, &d_in[i]));
        if (do_mo) RB_TRY(op.alloc(pc * slab_out, &d_mo[i]));
        if (do_mo && upper_out) RB_TRY(op.alloc(pc * ship_out, &d_pk[i]));
    }
    if (do_j) { RB_TRY(op.alloc(slab_in, &d_dm)); RB_TRY(op.alloc(nx, &d_d)); RB_TRY(op.alloc(slab_in, &d_j)); }
    if (do_k) { RB_TRY(op.alloc(nb * no, &d_ct)); RB_TRY(op.alloc(slab_in, &d_k)); }
    Pipe pipe;
    RB_TRY(pipe.init());
    if (do_mo) {
        RB_TRY(op.up(d_cl, cl, nb * nl));
        if (!same_c) RB_TRY(op.up(d_cr, cr, nb * nr));
    }
    if (do_j) RB_TRY(op.up(d_dm, dm, slab_in));
    if (do_k) RB_TRY(op.up(d_ct, ct, nb * no));
    const double t_alloc = now_ms();
    // (A ramp of small first/last chunks was measured twice and is slower -- 123.7 vs 118.8 ms at config C: the D2H rows
    //  are pn*8 bytes wide, so short chunks make the 2-D copy inefficient; 128- and 512-slab chunks: 147.6 / 140.2 ms.
    //  The pass is bound by PCIe duplex bandwidth, 9.8 GB at ~94 GB/s = 104 ms -- profiles/r01_e2e_variants.md.)
    // Chunk schedule: optional short head chunks (REST_B200_HEAD="64,128": the D2H stream, which bounds the pass, starts
    // earlier), then equal chunks of pc slabs.
    std::vector<i64> sizes;
    {
        i64 rem = nx;
        const char *e = getenv("REST_B200_HEAD");
        std::string head = e ? e : RB_DEFAULT_HEAD;
        size_t pos = 0;
        while (pos < head.size() && rem > 2 * pc) {
            size_t comma = head.find(',', pos);
            if (comma == std::string::npos) comma = head.size();
            i64 v = atoll(head.substr(pos, comma - pos).c_str());
            pos = comma + 1;
            if (v < 8 || v > pc) continue;
            sizes.push_back(v); rem -= v;
        }
        while (rem > 0) { const i64 pn = rem < pc ? rem : pc; sizes.push_back(pn); rem -= pn; }
    }
    // Pageable caller buffers (a plain Vec<f64>) are bounced through pinned blocks by host threads (see PinBlock above);
    // pinned / registered buffers go straight to the DMA engines.  REST_B200_BOUNCE=0 disables the bounce.
    bool bounce = true;
    if (const char *e = getenv("REST_B200_BOUNCE")) bounce = atoi(e) != 0;
    bool in_pageable = bounce && nb > 0 && !host_ptr_is_pinned(ri3ao);
    bool out_pageable = bounce && do_mo && !host_ptr_is_pinned(out);
    double *pin_in[2] = {nullptr, nullptr}, *pin_out[2] = {nullptr, nullptr};
    if (in_pageable)
        for (int i = 0; i < 2; ++i) if (!(pin_in[i] = (double *)pin_take(pc * slab_in * 8))) in_pageable = false;
    if (out_pageable)
        for (int i = 0; i < 2; ++i) if (!(pin_out[i] = (double *)pin_take(pc * ship_out * 8))) out_pageable = false;
    // REST_B200_TRACE: per-chunk timeline (timing events; start of H2D, end of H2D / compute / D2H relative to t0)
    std::vector<cudaEvent_t> tl;
    cudaEvent_t tl0 = nullptr;
    auto mark = [&](cudaStream_t st) {
        if (!trace) return;
        cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); tl.push_back(e);
    };
    if (trace) { cudaEventCreate(&tl0); cudaEventRecord(tl0, pipe.s_in); }
    int step = 0;
    i64 p0 = 0;
    for (size_t ci = 0; ci < sizes.size(); p0 += sizes[ci], ++ci, ++step) {
        const int s = step & 1;
        const i64 pn = sizes[ci];
        // H2D of this chunk may start once the compute that last read d_in[s] is done
        if (step >= 2) RB_CUDA(cudaStreamWaitEvent(pipe.s_in, pipe.comp_done[s], 0));
        const double *h_src = ri3ao + p0 * slab_in;
        if (in_pageable) { // the H2D that last read pin_in[s] (two chunks ago) is long done; the GPU is busy with chunk ci-1
            if (step >= 2) RB_CUDA(cudaEventSynchronize(pipe.in_done[s]));
            host_copy_2d(pin_in[s], pn * slab_in, h_src, pn * slab_in, pn * slab_in, 1);
            h_src = pin_in[s];
        }
        mark(pipe.s_in);
        if (nb > 0 && !symm_in)
            RB_CUDA(cudaMemcpyAsync(d_in[s], h_src, (size_t)(pn * slab_in) * 8, cudaMemcpyHostToDevice, pipe.s_in));
        if (nb > 0 && symm_in) {
            // rows 0 .. c0 + cw - 1 of the columns c0 .. c0 + cw - 1 of EVERY slab of the chunk: one pitched 3-D copy per block of
            // 32 columns (row pitch nb, slice pitch nb^2) -- 52-55 % of the bytes of the full slabs, ~19 copies per chunk at nb = 600
            const i64 CB = 32;
            for (i64 c0 = 0; c0 < nb; c0 += CB) {
                const i64 cw = nb - c0 < CB ? nb - c0 : CB, rows = c0 + cw;
                cudaMemcpy3DParms q = {};
                q.srcPtr = make_cudaPitchedPtr((void *)(h_src + c0 * nb), (size_t)nb * 8, (size_t)nb * 8, (size_t)nb);
                q.dstPtr = make_cudaPitchedPtr((void *)(d_in[s] + c0 * nb), (size_t)nb * 8, (size_t)nb * 8, (size_t)nb);
                q.extent = make_cudaExtent((size_t)rows * 8, (size_t)cw, (size_t)pn);
                q.kind = cudaMemcpyHostToDevice;
                RB_CUDA(cudaMemcpy3DAsync(&q, pipe.s_in));
            }
        }
        RB_CUDA(cudaEventRecord(pipe.in_done[s], pipe.s_in));
        mark(pipe.s_in);
        // compute needs the chunk in HBM and the previous D2H out of d_out[s]
        RB_CUDA(cudaStreamWaitEvent(ctx->stream, pipe.in_done[s], 0));
        if (step >= 2) RB_CUDA(cudaStreamWaitEvent(ctx->stream, pipe.out_done[s], 0));
        if (symm_in && nb > 1) {
            const i64 nt = rb_cdiv(nb, 32);
            rb_mirror_upper_slabs_kernel<<<dim3((unsigned)(nt * (nt + 1) / 2), (unsigned)pn), 256, 0, ctx->stream>>>(d_in[s], nb, slab_in);
            RB_LAUNCHED(ctx);
        }
        if (do_mo) RB_TRY(rb_ri_ao2mo(ctx, d_cl, nl, d_cr, nr, d_in[s], d_mo[s], nb_, (int)pn, pn));
        const double *d_ship = d_mo[s];
        if (do_mo && upper_out) {
            const int vec4 = ((pn & 3) == 0 && rb_aligned32(d_mo[s]) && rb_aligned32(d_pk[s])) ? 1 : 0;
            i64 bx = rb_cdiv(rb_cdiv((i64)nl * pn, vec4 ? 8 : 2), 256);
            if (bx > 32) bx = 32;
            if (bx < 1) bx = 1;
            rb_ri3mo_pack_pairs_kernel<<<dim3((unsigned)bx, (unsigned)nl), 256, 0, ctx->stream>>>(d_mo[s], d_pk[s], pn, nl, vec4);
            RB_LAUNCHED(ctx);
            d_ship = d_pk[s];
        }
        if (do_j && slab_in > 0) {
            RB_TRY(rb_ri_dp(ctx, d_in[s], d_dm, d_d + p0, nb_, (int)pn));
            RB_TRY(rb_dgemv(ctx, 'N', (int)slab_in, (int)pn, 1.0, d_in[s], slab_in, d_d + p0, 1, p0 == 0 ? 0.0 : 1.0, d_j, 1));
        }
        if (do_k && slab_in > 0) RB_TRY(rb_ri_k_upper(ctx, d_in[s], d_ct, no, d_k, nb, pn, p0 == 0 ? 0.0 : 1.0));
        RB_CUDA(cudaEventRecord(pipe.comp_done[s], ctx->stream));
        mark(ctx->stream);
        if (do_mo) {
            // D2H: rows of pn doubles into the P-fastest host tensor (pitch nx)
            RB_CUDA(cudaStreamWaitEvent(pipe.s_out, pipe.comp_done[s], 0));
            if (out_pageable) // dense copy into the pinned block; host threads scatter it one chunk later
                RB_CUDA(cudaMemcpyAsync(pin_out[s], d_ship, (size_t)(pn * ship_out) * 8, cudaMemcpyDeviceToHost, pipe.s_out));
            else
                RB_CUDA(cudaMemcpy2DAsync(out + p0, (size_t)nx * 8, d_ship, (size_t)pn * 8, (size_t)pn * 8, (size_t)ship_out,
                                          cudaMemcpyDeviceToHost, pipe.s_out));
            RB_CUDA(cudaEventRecord(pipe.out_done[s], pipe.s_out));
            mark(pipe.s_out);
            if (out_pageable && ci >= 1) { // previous chunk: its D2H ran while this chunk was copied in and enqueued
                const i64 pp = sizes[ci - 1];
                RB_CUDA(cudaEventSynchronize(pipe.out_done[s ^ 1]));
                host_copy_2d(out + (p0 - pp), nx, pin_out[s ^ 1], pp, pp, ship_out);
            }
        }
    }
    if (out_pageable && !sizes.empty()) { // last chunk
        const i64 pp = sizes.back();
        const int s = (step - 1) & 1;
        RB_CUDA(cudaEventSynchronize(pipe.out_done[s]));
        host_copy_2d(out + (nx - pp), nx, pin_out[s], pp, pp, ship_out);
    }
    const double t_enq = now_ms();
    if (do_k && slab_in > 0) RB_TRY(rb_symmetrize(ctx, d_k, nb, nb, true));
    if (do_j) {
        if (d_out) RB_TRY(op.down(d_out, d_d, nx));
        if (j_out) RB_TRY(op.down(j_out, d_j, slab_in));
    }
    if (do_k) RB_TRY(op.down(k_out, d_k, slab_in));
    RB_CUDA(cudaStreamSynchronize(pipe.s_out));
    RB_CUDA(cudaStreamSynchronize(pipe.s_in));
    RB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (trace && do_mo && tl.size() == 4 * sizes.size()) {
        for (size_t ci = 0; ci < sizes.size(); ++ci) {
            float t[4];
            for (int q = 0; q < 4; ++q) cudaEventElapsedTime(&t[q], tl0, tl[4 * ci + q]);
            fprintf(stderr, "[rest_b200]   chunk %2zu (%4lld slabs): H2D %7.2f -> %7.2f  compute done %7.2f  D2H done %7.2f ms\n", ci,
                    (long long)sizes[ci], t[0], t[1], t[2], t[3]);
        }
    }
    for (auto e : tl) cudaEventDestroy(e);
    if (tl0) cudaEventDestroy(tl0);
    if (trace)
        fprintf(stderr, "[rest_b200] ri stream: nb=%lld nx=%lld pc=%lld chunks=%d bounce(in,out)=%d,%d threads=%d  setup %.2f ms, enqueue %.2f ms, drain %.2f ms\n",
                (long long)nb, (long long)nx, (long long)pc, step, (int)in_pageable, (int)out_pageable, host_threads(),
                t_alloc - t_start, t_enq - t_alloc, now_ms() - t_enq);
    return RB_OK;
}

int host_ao2mo(const double *cl, int nl, const double *cr, int nr, const double *ri3ao, double *out, int nb, int nx)
{
    if (nx == 0 || nl == 0 || nr == 0) return RB_OK;
    return host_ri_stream(cl, nl, cr, nr, ri3ao, out, nb, nx, nullptr, nullptr, 0, nullptr, nullptr, nullptr);
}

int host_gemm_block(const double *a, i64 rows_a, i64 sra, i64 lra, i64 sca, i64 lca, char opa, const double *b,
                    i64 rows_b, i64 srb, i64 lrb, i64 scb, i64 lcb, char opb, double *c, i64 rows_c, i64 src, i64 lrc,
                    i64 scc, i64 lcc, double alpha, double beta)
{
    RB_REQUIRE((opa == 'N' || opa == 'T') && (opb == 'N' || opb == 'T'), "general_dgemm_f: op must be 'N' or 'T'");
    const i64 m = lrc, n = lcc, k = (opa == 'N') ? lca : lra;
    RB_REQUIRE(m >= 0 && n >= 0 && k >= 0, "general_dgemm_f: negative block length");
    RB_REQUIRE(((opa == 'N') ? lra : lca) == m, "general_dgemm_f: op(A) rows != rows of C block");
    RB_REQUIRE(((opb == 'N') ? lrb : lcb) == k && ((opb == 'N') ? lcb : lrb) == n, "general_dgemm_f: op(B) shape mismatch");
    if (m == 0 || n == 0) return RB_OK;
    HOST_CTX(op);
    double *da, *db, *dc;
    RB_TRY(op.alloc(lra * lca, &da));
    RB_TRY(op.alloc(lrb * lcb, &db));
    RB_TRY(op.alloc(m * n, &dc));
    RB_TRY(op.up2d(da, a + sra + sca * rows_a, lra, lca, rows_a));
    RB_TRY(op.up2d(db, b + srb + scb * rows_b, lrb, lcb, rows_b));
    if (beta != 0.0) RB_TRY(op.up2d(dc, c + src + scc * rows_c, m, n, rows_c));
    RB_TRY(rb_dgemm(op.ctx, opa, opb, (int)m, (int)n, (int)k, alpha, da, lra > 0 ? lra : 1, db, lrb > 0 ? lrb : 1, beta,
                    dc, m));
    RB_TRY(op.down2d(c + src + scc * rows_c, rows_c, dc, m, n));
    return op.sync();
}

} // namespace

// =====================================================================================================================
// (1) Fortran-ABI compat symbols
// =====================================================================================================================
extern "C" void ri_ao2mo_f_(const double *eigenvector, const double *ri3fn, double *ri3mo, const int *num_states,
                            const int *num_basis, const int *num_auxbas)
{
    die_if(host_ao2mo(eigenvector, *num_states, eigenvector, *num_states, ri3fn, ri3mo, *num_basis, *num_auxbas),
           "ri_ao2mo_f_");
}

extern "C" void general_dgemm_f_(const double *matr_a, const int *rows_a, const int *columns_a, const int *start_row_a,
                                 const int *len_row_a, const int *start_column_a, const int *len_column_a,
                                 const char *opa, const double *matr_b, const int *rows_b, const int *columns_b,
                                 const int *start_row_b, const int *len_row_b, const int *start_column_b,
                                 const int *len_column_b, const char *opb, double *matr_c, const int *rows_c,
                                 const int *columns_c, const int *start_row_c, const int *len_row_c,
                                 const int *start_column_c, const int *len_column_c, const double *alpha,
                                 const double *beta)
{
    (void)columns_a; (void)columns_b; (void)columns_c;
    die_if(host_gemm_block(matr_a, *rows_a, *start_row_a, *len_row_a, *start_column_a, *len_column_a, *opa, matr_b,
                           *rows_b, *start_row_b, *len_row_b, *start_column_b, *len_column_b, *opb, matr_c, *rows_c,
                           *start_row_c, *len_row_c, *start_column_c, *len_column_c, *alpha, *beta),
           "general_dgemm_f_");
}

static int host_special_dgemm(double *t, int x_a, int y_a, int z_a, int sx, int lx, int sz, int lz, const double *b,
                              int rows_b, int srb, int scb, int lcb, double alpha, double beta)
{
    RB_REQUIRE(x_a >= 0 && y_a >= 0 && z_a >= 0, "special_dgemm_f_01: negative dimension");
    const i64 total = (i64)x_a * y_a * z_a;
    if (total == 0 || lx == 0 || lz == 0) return RB_OK;
    HOST_CTX(op);
    double *dt, *db;
    RB_TRY(op.alloc(total, &dt));
    RB_TRY(op.alloc((i64)lz * lcb, &db));
    RB_TRY(op.up(dt, t, total));
    RB_TRY(op.up2d(db, b + srb + (i64)scb * rows_b, lz, lcb, rows_b));
    RB_TRY(rb_special_dgemm_01(op.ctx, dt, x_a, y_a, z_a, sx, lx, sz, lz, db, lz, lcb, alpha, beta));
    RB_TRY(op.down(t, dt, total));
    return op.sync();
}

extern "C" void special_dgemm_f_01_(double *ten3_a, const int *x_a, const int *y_a, const int *z_a,
                                    const int *start_x_a, const int *len_x_a, const int *i_y, const int *start_z_a,
                                    const int *len_z_a, const double *matr_b, const int *rows_b, const int *columns_b,
                                    const int *start_row_b, const int *len_row_b, const int *start_column_b,
                                    const int *len_column_b, const double *alpha, const double *beta)
{
    (void)i_y; (void)columns_b; (void)len_row_b; // i_y_a is unused by the Fortran as well (restmatr.f90:111-154)
    die_if(host_special_dgemm(ten3_a, *x_a, *y_a, *z_a, *start_x_a, *len_x_a, *start_z_a, *len_z_a, matr_b, *rows_b,
                              *start_row_b, *start_column_b, *len_column_b, *alpha, *beta),
           "special_dgemm_f_01_");
}

// copy_* : upload the source box, run the strided-copy kernel into a device image of the destination box,
// download the box.  (Only the destination BOX is written back, so untouched elements keep their host values.)
static int host_copy_box(const double *f, i64 f0, i64 fs1, i64 fs2, i64 fs3, double *t, i64 t0, i64 ts1, i64 ts2,
                         i64 ts3, i64 n1, i64 n2, i64 n3)
{
    if (n1 <= 0 || n2 <= 0 || n3 <= 0) return RB_OK;
    HOST_CTX(op);
    // source box -> dense device buffer [n1,n2,n3]
    double *ds, *dd;
    RB_TRY(op.alloc(n1 * n2 * n3, &ds));
    RB_TRY(op.alloc(n1 * n2 * n3, &dd));
    // gather rows of the source box (unit stride along whichever axis has stride 1; generic fallback: element rows)
    if (fs1 == 1) {
        for (i64 k = 0; k < n3; ++k)
            RB_CUDA(cudaMemcpy2DAsync(ds + k * n1 * n2, (size_t)n1 * 8, f + f0 + k * fs3, (size_t)fs2 * 8, (size_t)n1 * 8,
                                      (size_t)n2, cudaMemcpyHostToDevice, op.ctx->stream));
    } else if (fs2 % fs1 == 0 && fs2 / fs1 >= n1) {
        // x3-fixed plane of an RI tensor (mode 2): every element sits alone, fs1 doubles from the next one along x1 and
        // fs2 = (extent of x1) * fs1 along x2.  Seen as a pitched volume of 8-byte rows (pitch fs1*8, fs2/fs1 rows per
        // slice) the whole plane is ONE 3-D copy per k instead of one 2-D copy of 8-byte rows per (j, k).
        for (i64 k = 0; k < n3; ++k) {
            cudaMemcpy3DParms q = {};
            q.srcPtr = make_cudaPitchedPtr((void *)(f + f0 + k * fs3), (size_t)fs1 * 8, 8, (size_t)(fs2 / fs1));
            q.dstPtr = make_cudaPitchedPtr((void *)(ds + k * n1 * n2), 8, 8, (size_t)n1);
            q.extent = make_cudaExtent(8, (size_t)n1, (size_t)n2);
            q.kind = cudaMemcpyHostToDevice;
            RB_CUDA(cudaMemcpy3DAsync(&q, op.ctx->stream));
        }
    } else {
        for (i64 k = 0; k < n3; ++k)
            for (i64 j = 0; j < n2; ++j)
                RB_CUDA(cudaMemcpy2DAsync(ds + (j + k * n2) * n1, 8, f + f0 + j * fs2 + k * fs3, (size_t)fs1 * 8, 8,
                                          (size_t)n1, cudaMemcpyHostToDevice, op.ctx->stream));
    }
    RB_TRY(rb_copy3d(op.ctx, ds, 0, 1, n1, n1 * n2, dd, 0, 1, n1, n1 * n2, n1, n2, n3));
    if (ts1 == 1) {
        for (i64 k = 0; k < n3; ++k)
            RB_CUDA(cudaMemcpy2DAsync(t + t0 + k * ts3, (size_t)ts2 * 8, dd + k * n1 * n2, (size_t)n1 * 8, (size_t)n1 * 8,
                                      (size_t)n2, cudaMemcpyDeviceToHost, op.ctx->stream));
    } else if (ts2 % ts1 == 0 && ts2 / ts1 >= n1) {
        for (i64 k = 0; k < n3; ++k) {
            cudaMemcpy3DParms q = {};
            q.srcPtr = make_cudaPitchedPtr((void *)(dd + k * n1 * n2), 8, 8, (size_t)n1);
            q.dstPtr = make_cudaPitchedPtr((void *)(t + t0 + k * ts3), (size_t)ts1 * 8, 8, (size_t)(ts2 / ts1));
            q.extent = make_cudaExtent(8, (size_t)n1, (size_t)n2);
            q.kind = cudaMemcpyDeviceToHost;
            RB_CUDA(cudaMemcpy3DAsync(&q, op.ctx->stream));
        }
    } else {
        for (i64 k = 0; k < n3; ++k)
            for (i64 j = 0; j < n2; ++j)
                RB_CUDA(cudaMemcpy2DAsync(t + t0 + j * ts2 + k * ts3, (size_t)ts1 * 8, dd + (j + k * n2) * n1, 8, 8,
                                          (size_t)n1, cudaMemcpyDeviceToHost, op.ctx->stream));
    }
    return op.sync();
}

static bool in_box(i64 s, i64 l, i64 d) { return s >= 0 && l >= 0 && s + l <= d; }

static int ri_plane(int mod, i64 X, i64 Y, i64 Z, i64 s1, i64 s2, i64 x3, i64 l1, i64 l2, i64 *off, i64 *st1, i64 *st2)
{
    if (mod == 0) {
        RB_REQUIRE(in_box(s1, l1, X) && in_box(s2, l2, Y) && x3 >= 0 && x3 < Z, "copy: block outside tensor");
        *off = s1 + s2 * X + x3 * X * Y; *st1 = 1; *st2 = X;
    } else if (mod == 1) {
        RB_REQUIRE(in_box(s1, l1, X) && in_box(s2, l2, Z) && x3 >= 0 && x3 < Y, "copy: block outside tensor");
        *off = s1 + x3 * X + s2 * X * Y; *st1 = 1; *st2 = X * Y;
    } else {
        RB_REQUIRE(in_box(s1, l1, Y) && in_box(s2, l2, Z) && x3 >= 0 && x3 < X, "copy: block outside tensor");
        *off = x3 + s1 * X + s2 * X * Y; *st1 = X; *st2 = X * Y;
    }
    return RB_OK;
}

extern "C" void copy_mm_(const int *x_len, const int *y_len, const double *f_matr, const int *f_x_len,
                         const int *f_y_len, const int *f_x_start, const int *f_y_start, double *t_matr,
                         const int *t_x_len, const int *t_y_len, const int *t_x_start, const int *t_y_start)
{
    int st = RB_OK;
    if (!(in_box(*f_x_start, *x_len, *f_x_len) && in_box(*f_y_start, *y_len, *f_y_len) &&
          in_box(*t_x_start, *x_len, *t_x_len) && in_box(*t_y_start, *y_len, *t_y_len))) {
        rb_set_error("copy_mm_: block outside matrix");
        st = RB_ERR_INVALID;
    } else {
        st = host_copy_box(f_matr, *f_x_start + (i64)*f_y_start * *f_x_len, 1, *f_x_len, 0, t_matr,
                           *t_x_start + (i64)*t_y_start * *t_x_len, 1, *t_x_len, 0, *x_len, *y_len, 1);
    }
    die_if(st, "copy_mm_");
}

extern "C" void copy_mr_(const int *x_len, const int *y_len, const double *f_matr, const int *f_x_len,
                         const int *f_y_len, const int *f_x_start, const int *f_y_start, double *t_ri,
                         const int *t_x_len, const int *t_y_len, const int *t_z_len, const int *t_x_start,
                         const int *t_y_start, const int *t_x3, const int *t_mod)
{
    if (*t_mod < 0 || *t_mod > 2) return; // restmatr.f90:227-237
    int st;
    i64 off, s1, s2;
    if (!(in_box(*f_x_start, *x_len, *f_x_len) && in_box(*f_y_start, *y_len, *f_y_len))) {
        rb_set_error("copy_mr_: block outside matrix");
        st = RB_ERR_INVALID;
    } else if ((st = ri_plane(*t_mod, *t_x_len, *t_y_len, *t_z_len, *t_x_start, *t_y_start, *t_x3, *x_len, *y_len, &off,
                              &s1, &s2)) == RB_OK) {
        st = host_copy_box(f_matr, *f_x_start + (i64)*f_y_start * *f_x_len, 1, *f_x_len, 0, t_ri, off, s1, s2, 0, *x_len,
                           *y_len, 1);
    }
    die_if(st, "copy_mr_");
}

extern "C" void copy_rm_(const int *x_len, const int *y_len, const double *f_ri, const int *f_x_len, const int *f_y_len,
                         const int *f_z_len, const int *f_x_start, const int *f_y_start, const int *f_x3,
                         const int *f_mod, double *t_matr, const int *t_x_len, const int *t_y_len, const int *t_x_start,
                         const int *t_y_start)
{
    if (*f_mod < 0 || *f_mod > 2) return;
    int st;
    i64 off, s1, s2;
    if (!(in_box(*t_x_start, *x_len, *t_x_len) && in_box(*t_y_start, *y_len, *t_y_len))) {
        rb_set_error("copy_rm_: block outside matrix");
        st = RB_ERR_INVALID;
    } else if ((st = ri_plane(*f_mod, *f_x_len, *f_y_len, *f_z_len, *f_x_start, *f_y_start, *f_x3, *x_len, *y_len, &off,
                              &s1, &s2)) == RB_OK) {
        st = host_copy_box(f_ri, off, s1, s2, 0, t_matr, *t_x_start + (i64)*t_y_start * *t_x_len, 1, *t_x_len, 0, *x_len,
                           *y_len, 1);
    }
    die_if(st, "copy_rm_");
}

extern "C" void copy_rr_(const int *x_len, const int *y_len, const int *z_len, const double *f_ri, const int *f_x_len,
                         const int *f_y_len, const int *f_z_len, const int *f_x_start, const int *f_y_start,
                         const int *f_z_start, double *t_ri, const int *t_x_len, const int *t_y_len, const int *t_z_len,
                         const int *t_x_start, const int *t_y_start, const int *t_z_start)
{
    int st;
    if (!(in_box(*f_x_start, *x_len, *f_x_len) && in_box(*f_y_start, *y_len, *f_y_len) &&
          in_box(*f_z_start, *z_len, *f_z_len) && in_box(*t_x_start, *x_len, *t_x_len) &&
          in_box(*t_y_start, *y_len, *t_y_len) && in_box(*t_z_start, *z_len, *t_z_len))) {
        rb_set_error("copy_rr_: box outside tensor");
        st = RB_ERR_INVALID;
    } else {
        const i64 FX = *f_x_len, FY = *f_y_len, TX = *t_x_len, TY = *t_y_len;
        st = host_copy_box(f_ri, *f_x_start + *f_y_start * FX + *f_z_start * FX * FY, 1, FX, FX * FY, t_ri,
                           *t_x_start + *t_y_start * TX + *t_z_start * TX * TY, 1, TX, TX * TY, *x_len, *y_len, *z_len);
    }
    die_if(st, "copy_rr_");
}

// =====================================================================================================================
// (2) rb_host_* wrappers
// =====================================================================================================================
extern "C" int rb_host_ri_ao2mo(const double *c_left, int nl, const double *c_right, int nr, const double *ri3ao,
                                double *out, int nb, int nx)
{
    return host_ao2mo(c_left, nl, c_right, nr, ri3ao, out, nb, nx);
}

extern "C" int rb_host_ri_ao2mo_jk(const double *c_left, int nl, const double *c_right, int nr, const double *ri3ao,
                                   double *ri3mo, int nb, int nx, const double *dm, const double *ct, int no, double *d,
                                   double *j, double *k)
{
    return host_ri_stream(c_left, nl, c_right, nr, ri3ao, ri3mo, nb, nx, dm, ct, no, d, j, k);
}

extern "C" int rb_host_ri_ao2mo_jk_upper(const double *c, int nmo, const double *ri3ao, double *ri3mo_upper, int nb, int nx,
                                         const double *dm, const double *ct, int no, double *d, double *j, double *k)
{
    return host_ri_stream(c, nmo, c, nmo, ri3ao, ri3mo_upper, nb, nx, dm, ct, no, d, j, k, true);
}

extern "C" int rb_host_ri_ao2mo_jk_symm(const double *c, int nmo, const double *ri3ao, double *ri3mo_upper, int nb, int nx,
                                        const double *dm, const double *ct, int no, double *d, double *j, double *k)
{
    return host_ri_stream(c, nmo, c, nmo, ri3ao, ri3mo_upper, nb, nx, dm, ct, no, d, j, k, true, true);
}

extern "C" int rb_host_dgemm(char ta, char tb, int m, int n, int k, double alpha, const double *a, int lda,
                             const double *b, int ldb, double beta, double *c, int ldc)
{
    RB_REQUIRE((rb_is_n(ta) || rb_is_t(ta)) && (rb_is_n(tb) || rb_is_t(tb)), "rb_host_dgemm: trans must be 'N' or 'T'");
    RB_REQUIRE(m >= 0 && n >= 0 && k >= 0, "rb_host_dgemm: negative dimension");
    if (m == 0 || n == 0) return RB_OK;
    const i64 ra = rb_is_n(ta) ? m : k, ca = rb_is_n(ta) ? k : m;
    const i64 rb_ = rb_is_n(tb) ? k : n, cb = rb_is_n(tb) ? n : k;
    RB_REQUIRE(lda >= (ra > 1 ? ra : 1) && ldb >= (rb_ > 1 ? rb_ : 1) && ldc >= m, "rb_host_dgemm: leading dimension too small");
    HOST_CTX(op);
    double *da, *db, *dc;
    RB_TRY(op.alloc(ra * ca, &da));
    RB_TRY(op.alloc(rb_ * cb, &db));
    RB_TRY(op.alloc((i64)m * n, &dc));
    RB_TRY(op.up2d(da, a, ra, ca, lda));
    RB_TRY(op.up2d(db, b, rb_, cb, ldb));
    if (beta != 0.0) RB_TRY(op.up2d(dc, c, m, n, ldc));
    RB_TRY(rb_dgemm(op.ctx, ta, tb, m, n, k, alpha, da, ra > 0 ? ra : 1, db, rb_ > 0 ? rb_ : 1, beta, dc, m));
    RB_TRY(op.down2d(c, ldc, dc, m, n));
    return op.sync();
}

extern "C" int rb_host_dsyrk(char uplo, char trans, int n, int k, double alpha, const double *a, int lda, double beta,
                             double *c, int ldc)
{
    RB_REQUIRE(rb_is_u(uplo) || rb_is_l(uplo), "rb_host_dsyrk: uplo must be 'U' or 'L'");
    RB_REQUIRE(rb_is_n(trans) || rb_is_t(trans), "rb_host_dsyrk: trans must be 'N' or 'T'");
    RB_REQUIRE(n >= 0 && k >= 0, "rb_host_dsyrk: negative dimension");
    if (n == 0) return RB_OK;
    const i64 ra = rb_is_n(trans) ? n : k, ca = rb_is_n(trans) ? k : n;
    RB_REQUIRE(lda >= (ra > 1 ? ra : 1) && ldc >= n, "rb_host_dsyrk: leading dimension too small");
    HOST_CTX(op);
    double *da, *dc;
    RB_TRY(op.alloc(ra * ca, &da));
    RB_TRY(op.alloc((i64)n * n, &dc));
    RB_TRY(op.up2d(da, a, ra, ca, lda));
    // C travels both ways in full: the kernel touches only the `uplo` triangle, so the other one round-trips unchanged.
    RB_TRY(op.up2d(dc, c, n, n, ldc));
    RB_TRY(rb_dsyrk(op.ctx, uplo, trans, n, k, alpha, da, ra > 0 ? ra : 1, beta, dc, n));
    RB_TRY(op.down2d(c, ldc, dc, n, n));
    return op.sync();
}

extern "C" int rb_host_dgemv(char trans, int m, int n, double alpha, const double *a, int lda, const double *x,
                             int incx, double beta, double *y, int incy)
{
    RB_REQUIRE(rb_is_n(trans) || rb_is_t(trans), "rb_host_dgemv: trans must be 'N' or 'T'");
    RB_REQUIRE(m >= 0 && n >= 0 && incx != 0 && incy != 0, "rb_host_dgemv: bad dimension / increment");
    if (m == 0 || n == 0) return RB_OK;
    RB_REQUIRE(lda >= m, "rb_host_dgemv: lda too small");
    const i64 lenx = rb_is_n(trans) ? n : m, leny = rb_is_n(trans) ? m : n;
    const i64 ax = incx > 0 ? incx : -incx, ay = incy > 0 ? incy : -incy;
    const i64 nxv = 1 + (lenx - 1) * ax, nyv = 1 + (leny - 1) * ay;
    HOST_CTX(op);
    double *da, *dx, *dy;
    RB_TRY(op.alloc((i64)m * n, &da));
    RB_TRY(op.alloc(nxv, &dx));
    RB_TRY(op.alloc(nyv, &dy));
    RB_TRY(op.up2d(da, a, m, n, lda));
    RB_TRY(op.up(dx, x, nxv));
    RB_TRY(op.up(dy, y, nyv));
    RB_TRY(rb_dgemv(op.ctx, trans, m, n, alpha, da, m, dx, incx, beta, dy, incy));
    RB_TRY(op.down(y, dy, nyv));
    return op.sync();
}

extern "C" int rb_host_dsymm(char side, char uplo, int m, int n, double alpha, const double *a, int lda,
                             const double *b, int ldb, double beta, double *c, int ldc)
{
    RB_REQUIRE(rb_is_l(side) || side == 'R' || side == 'r', "rb_host_dsymm: side must be 'L' or 'R'");
    RB_REQUIRE(rb_is_u(uplo) || rb_is_l(uplo), "rb_host_dsymm: uplo must be 'U' or 'L'");
    RB_REQUIRE(m >= 0 && n >= 0, "rb_host_dsymm: negative dimension");
    if (m == 0 || n == 0) return RB_OK;
    const i64 ka = rb_is_l(side) ? m : n;
    RB_REQUIRE(lda >= ka && ldb >= m && ldc >= m, "rb_host_dsymm: leading dimension too small");
    HOST_CTX(op);
    double *da, *db, *dc;
    RB_TRY(op.alloc(ka * ka, &da));
    RB_TRY(op.alloc((i64)m * n, &db));
    RB_TRY(op.alloc((i64)m * n, &dc));
    RB_TRY(op.up2d(da, a, ka, ka, lda));
    RB_TRY(op.up2d(db, b, m, n, ldb));
    if (beta != 0.0) RB_TRY(op.up2d(dc, c, m, n, ldc));
    RB_TRY(rb_dsymm(op.ctx, side, uplo, m, n, alpha, da, ka, db, m, beta, dc, m));
    RB_TRY(op.down2d(c, ldc, dc, m, n));
    return op.sync();
}

extern "C" int rb_host_to_matrixupper(const double *full, int64_t n, double *packed)
{
    RB_REQUIRE(n >= 0, "rb_host_to_matrixupper: negative n");
    if (n == 0) return RB_OK;
    HOST_CTX(op);
    double *df, *dp;
    const i64 np = n * (n + 1) / 2;
    RB_TRY(op.alloc(n * n, &df));
    RB_TRY(op.alloc(np, &dp));
    RB_TRY(op.up(df, full, n * n));
    RB_TRY(rb_pack_upper(op.ctx, df, n, dp));
    RB_TRY(op.down(packed, dp, np));
    return op.sync();
}

// exact integer inverse of len = n(n+1)/2 (the reference uses an f64 sqrt, matrixupper.rs:331-332; identical for
// every representable triangular len)
static i64 tri_dim(i64 len)
{
    if (len <= 0) return 0;
    i64 n = (i64)((sqrt(1.0 + 8.0 * (double)len) - 1.0) * 0.5);
    while (n * (n + 1) / 2 > len) --n;
    while ((n + 1) * (n + 2) / 2 <= len) ++n;
    return (n * (n + 1) / 2 == len) ? n : -1;
}

extern "C" int rb_host_to_matrixfull(const double *packed, int64_t len, double *full)
{
    RB_REQUIRE(len >= 0, "rb_host_to_matrixfull: negative length");
    if (len == 0) return RB_OK; // MatrixFull::empty()
    const i64 n = tri_dim(len);
    RB_REQUIRE(n > 0, "rb_host_to_matrixfull: length %lld is not n(n+1)/2 (the reference returns None)", (long long)len);
    HOST_CTX(op);
    double *df, *dp;
    RB_TRY(op.alloc(n * n, &df));
    RB_TRY(op.alloc(len, &dp));
    RB_TRY(op.up(dp, packed, len));
    RB_TRY(rb_unpack_upper(op.ctx, dp, n, df));
    RB_TRY(op.down(full, df, n * n));
    return op.sync();
}

extern "C" int rb_host_ri_pack_symm(const double *ri, int64_t nao, int64_t naux, double *out)
{
    RB_REQUIRE(nao >= 0 && naux >= 0, "rb_host_ri_pack_symm: negative dimension");
    if (nao == 0 || naux == 0) return RB_OK;
    HOST_CTX(op);
    double *di, *dout;
    const i64 np = nao * (nao + 1) / 2;
    RB_TRY(op.alloc(nao * nao * naux, &di));
    RB_TRY(op.alloc(np * naux, &dout));
    RB_TRY(op.up(di, ri, nao * nao * naux));
    RB_TRY(rb_ri_pack_symm(op.ctx, di, nao, naux, dout));
    RB_TRY(op.down(out, dout, np * naux));
    return op.sync();
}

extern "C" int rb_host_ri_transpose(const double *in, int64_t i, int64_t j, int64_t k, int which, double *out)
{
    RB_REQUIRE(i >= 0 && j >= 0 && k >= 0, "rb_host_ri_transpose: negative dimension");
    RB_REQUIRE(which >= 0 && which <= 3, "rb_host_ri_transpose: which must be 0..3");
    const i64 n = i * j * k;
    if (n == 0) return RB_OK;
    HOST_CTX(op);
    double *di, *dout;
    RB_TRY(op.alloc(n, &di));
    RB_TRY(op.alloc(n, &dout));
    RB_TRY(op.up(di, in, n));
    RB_TRY(rb_ri_transpose(op.ctx, di, i, j, k, which, dout));
    RB_TRY(op.down(out, dout, n));
    return op.sync();
}

extern "C" int rb_host_matrix_transpose(const double *in, int64_t rows, int64_t cols, double *out)
{
    RB_REQUIRE(rows >= 0 && cols >= 0, "rb_host_matrix_transpose: negative dimension");
    const i64 n = rows * cols;
    if (n == 0) return RB_OK;
    HOST_CTX(op);
    double *di, *dout;
    RB_TRY(op.alloc(n, &di));
    RB_TRY(op.alloc(n, &dout));
    RB_TRY(op.up(di, in, n));
    RB_TRY(rb_matrix_transpose(op.ctx, di, rows, cols, dout));
    RB_TRY(op.down(out, dout, n));
    return op.sync();
}

extern "C" int rb_host_axpy(int opc, double *c, const double *p, double a, double b, int64_t n)
{
    RB_REQUIRE(opc >= 0 && opc <= 4, "rb_host_axpy: op must be 0..4");
    RB_REQUIRE(n >= 0, "rb_host_axpy: negative length");
    if (n == 0) return RB_OK;
    HOST_CTX(op);
    double *dc, *dp = nullptr;
    RB_TRY(op.alloc(n, &dc));
    RB_TRY(op.up(dc, c, n));
    if (opc != 2) { RB_TRY(op.alloc(n, &dp)); RB_TRY(op.up(dp, p, n)); }
    switch (opc) {
    case 0: RB_TRY(rb_self_scaled_add(op.ctx, dc, dp, b, n)); break;
    case 1: RB_TRY(rb_self_general_add(op.ctx, dc, dp, a, b, n)); break;
    case 2: RB_TRY(rb_self_multiple(op.ctx, dc, a, n)); break;
    case 3: RB_TRY(rb_self_add(op.ctx, dc, dp, n)); break;
    default: RB_TRY(rb_self_sub(op.ctx, dc, dp, n)); break;
    }
    RB_TRY(op.down(c, dc, n));
    return op.sync();
}

extern "C" int rb_host_einsum(int which, const double *a, const double *b, double *out, int64_t ni, int64_t nj)
{
    RB_REQUIRE(which >= 1 && which <= 3, "rb_host_einsum: which must be 1..3");
    RB_REQUIRE(ni >= 0 && nj >= 0, "rb_host_einsum: negative dimension");
    const i64 na = which == 3 ? ni : ni * nj, nb_ = which == 2 ? ni * nj : nj, no = which == 2 ? nj : ni * nj;
    if (no == 0) return RB_OK;
    HOST_CTX(op);
    double *da, *db, *dout;
    RB_TRY(op.alloc(na, &da));
    RB_TRY(op.alloc(nb_, &db));
    RB_TRY(op.alloc(no, &dout));
    RB_TRY(op.up(da, a, na));
    RB_TRY(op.up(db, b, nb_));
    if (which == 1) RB_TRY(rb_einsum_ij_j(op.ctx, da, ni, db, dout, ni, ni, nj));
    else if (which == 2) RB_TRY(rb_einsum_ip_ip(op.ctx, da, ni > 0 ? ni : 1, db, ni > 0 ? ni : 1, dout, ni, nj));
    else RB_TRY(rb_einsum_i_j(op.ctx, da, db, dout, ni, nj));
    RB_TRY(op.down(out, dout, no));
    return op.sync();
}

// ERIFold4 chunk scatter on a HOST tensor: the block only touches the rows [row_lo, row_hi] of the columns (k, l), k <= l, so
// just that window of every touched column crosses PCIe (up, scatter on the device, back); the rest of the tensor stays put.
extern "C" int rb_host_erifold4_chunk_copy(double *eri, int64_t size0, int64_t size1, int i0, int li, int j0, int lj, int k0, int lk,
                                           int l0, int ll, const double *buf, int mode)
{
    RB_REQUIRE(mode == 0 || mode == 1, "rb_host_erifold4_chunk_copy: mode must be 0 or 1");
    RB_REQUIRE(size0 >= 0 && size1 >= 0, "rb_host_erifold4_chunk_copy: bad tensor shape");
    RB_REQUIRE(i0 >= 0 && j0 >= 0 && k0 >= 0 && l0 >= 0 && li >= 0 && lj >= 0 && lk >= 0 && ll >= 0,
               "rb_host_erifold4_chunk_copy: negative range");
    const i64 total = (i64)li * lj * lk * ll;
    if (total == 0 || (mode == 1 && i0 > j0)) return RB_OK;
    RB_REQUIRE(eri && buf, "rb_host_erifold4_chunk_copy: NULL buffer");
    const i64 jlo = j0, jhi = (i64)j0 + lj - 1;
    i64 ihi = (i64)i0 + li - 1;
    if ((mode == 0 || i0 == j0) && ihi > jhi) ihi = jhi;
    const i64 row_lo = jlo * (jlo + 1) / 2 + i0, row_hi = jhi * (jhi + 1) / 2 + ihi;
    if (mode == 0 && i0 > jhi) return RB_OK; // no i <= j in the block
    RB_REQUIRE(row_hi < size0, "rb_host_erifold4_chunk_copy: the block reaches outside the folded tensor (row %lld of %lld)",
               (long long)row_hi, (long long)size0);
    const i64 rows = row_hi - row_lo + 1;
    HOST_CTX(op);
    double *dbuf, *dwin;
    RB_TRY(op.alloc(total, &dbuf));
    RB_TRY(op.up(dbuf, buf, total));
    // per l: the columns (k, l), k = k0 .. min(k0 + lk - 1, l), are consecutive: one pitched window [rows, nk]
    for (i64 lx = 0; lx < ll; ++lx) {
        const i64 l = (i64)l0 + lx;
        i64 khi = (i64)k0 + lk - 1;
        if (khi > l) khi = l;
        if (khi < k0) continue;
        const i64 nk = khi - k0 + 1, col0 = l * (l + 1) / 2 + k0;
        RB_REQUIRE(col0 + nk <= size1, "rb_host_erifold4_chunk_copy: the block reaches outside the folded tensor (column %lld of %lld)",
                   (long long)(col0 + nk - 1), (long long)size1);
        RB_TRY(op.alloc(rows * nk, &dwin));
        double *hwin = eri + col0 * size0 + row_lo;
        RB_TRY(op.up2d(dwin, hwin, rows, nk, size0));
        // the window is a [rows, nk] matrix whose element (row - row_lo, k - k0) is tensor element (row, (k, l)): scatter with the
        // tensor's index arithmetic by pointing the kernel at a virtual origin
        double *origin = dwin - (col0 * rows + row_lo);
        RB_TRY(rb_erifold4_scatter(op.ctx, origin, rows, dbuf + lx * (i64)li * lj * lk, i0, li, j0, lj, k0, lk, l, 1, mode));
        RB_TRY(op.down2d(hwin, size0, dwin, rows, nk));
    }
    return op.sync();
}

extern "C" int rb_host_ri_dp(const double *ri3ao, const double *dm, double *d, int nb, int nx)
{
    RB_REQUIRE(nb >= 0 && nx >= 0, "rb_host_ri_dp: negative dimension");
    if (nx == 0) return RB_OK;
    HOST_CTX(op);
    const i64 n2 = (i64)nb * nb;
    double *da, *dd, *dout;
    RB_TRY(op.alloc(n2 * nx, &da));
    RB_TRY(op.alloc(n2, &dd));
    RB_TRY(op.alloc(nx, &dout));
    RB_TRY(op.up(da, ri3ao, n2 * nx));
    RB_TRY(op.up(dd, dm, n2));
    RB_TRY(rb_ri_dp(op.ctx, da, dd, dout, nb, nx));
    RB_TRY(op.down(d, dout, nx));
    return op.sync();
}

extern "C" int rb_host_ri_j(const double *ri3ao, const double *d, double *j, int nb, int nx)
{
    RB_REQUIRE(nb >= 0 && nx >= 0, "rb_host_ri_j: negative dimension");
    const i64 n2 = (i64)nb * nb;
    if (n2 == 0) return RB_OK;
    HOST_CTX(op);
    double *da, *dd, *dj;
    RB_TRY(op.alloc(n2 * nx, &da));
    RB_TRY(op.alloc(nx, &dd));
    RB_TRY(op.alloc(n2, &dj));
    RB_TRY(op.up(da, ri3ao, n2 * nx));
    RB_TRY(op.up(dd, d, nx));
    RB_TRY(rb_ri_j(op.ctx, da, dd, dj, nb, nx));
    RB_TRY(op.down(j, dj, n2));
    return op.sync();
}

extern "C" int rb_host_ri_k(const double *ri3ao, const double *ct, int no, double *k, int nb, int nx)
{
    RB_REQUIRE(nb >= 0 && nx >= 0 && no >= 0, "rb_host_ri_k: negative dimension");
    const i64 n2 = (i64)nb * nb;
    if (n2 == 0) return RB_OK;
    HOST_CTX(op);
    double *da, *dc, *dk;
    RB_TRY(op.alloc(n2 * nx, &da));
    RB_TRY(op.alloc((i64)nb * no, &dc));
    RB_TRY(op.alloc(n2, &dk));
    RB_TRY(op.up(da, ri3ao, n2 * nx));
    RB_TRY(op.up(dc, ct, (i64)nb * no));
    RB_TRY(rb_ri_k(op.ctx, da, dc, no, dk, nb, nx));
    RB_TRY(op.down(k, dk, n2));
    return op.sync();
}

// (ia|jb)-type blocks of dense host tensors moA[np, nl_a, nr_a], moB[np, nl_b, nr_b] (moB may be
// moA): only the r-ranges the two boxes touch are uploaded (each is one contiguous block of a P-fastest tensor);
// out is the dense [lla*rla, llb*rlb] block.
extern "C" int rb_host_ri_iajb(int np, const double *mo_a, int nl_a, int nr_a, int l0a, int lla, int r0a, int rla,
                               const double *mo_b, int nl_b, int nr_b, int l0b, int llb, int r0b, int rlb, double *out)
{
    RB_REQUIRE(np >= 0 && nl_a >= 0 && nr_a >= 0 && nl_b >= 0 && nr_b >= 0, "rb_host_ri_iajb: negative dimension");
    RB_REQUIRE(l0a >= 0 && lla >= 0 && l0a + (i64)lla <= nl_a && r0a >= 0 && rla >= 0 && r0a + (i64)rla <= nr_a &&
                   l0b >= 0 && llb >= 0 && l0b + (i64)llb <= nl_b && r0b >= 0 && rlb >= 0 && r0b + (i64)rlb <= nr_b,
               "rb_host_ri_iajb: box outside the tensor");
    const i64 m = (i64)lla * rla, n = (i64)llb * rlb;
    if (m == 0 || n == 0) return RB_OK;
    RB_REQUIRE(out && (np == 0 || (mo_a && mo_b)), "rb_host_ri_iajb: NULL buffer");
    HOST_CTX(op);
    const i64 plane_a = (i64)np * nl_a, plane_b = (i64)np * nl_b;
    const bool same = mo_a == mo_b && nl_a == nl_b && r0a == r0b && rla == rlb; // one upload serves both sides
    double *da, *db, *dout;
    RB_TRY(op.alloc(plane_a * rla, &da));
    RB_TRY(op.up(da, mo_a + plane_a * r0a, plane_a * rla));
    if (same) db = da;
    else {
        RB_TRY(op.alloc(plane_b * rlb, &db));
        RB_TRY(op.up(db, mo_b + plane_b * r0b, plane_b * rlb));
    }
    RB_TRY(op.alloc(m * n, &dout));
    RB_TRY(rb_ri_iajb(op.ctx, np, da, np, nl_a, rla, l0a, lla, 0, rla, db, np, nl_b, rlb, l0b, llb, 0, rlb, 0.0, dout, m));
    RB_TRY(op.down(out, dout, m * n));
    return op.sync();
}

// RPA-type consumer on a dense host ri3mo[np, nl, nr]: out[P,Q] = sum_{(l,r) in box} w[l,r] mo[P,l,r] mo[Q,l,r]
// (w == NULL: all ones); out is the dense symmetric [np, np] matrix, overwritten.  Only the box's r-range is uploaded.
extern "C" int rb_host_ri_mo_pq(const double *mo, int np, int nl, int nr, int l0, int ll, int r0, int rl, const double *w,
                                double *out)
{
    RB_REQUIRE(np >= 0 && nl >= 0 && nr >= 0, "rb_host_ri_mo_pq: negative dimension");
    RB_REQUIRE(l0 >= 0 && ll >= 0 && l0 + (i64)ll <= nl && r0 >= 0 && rl >= 0 && r0 + (i64)rl <= nr,
               "rb_host_ri_mo_pq: box outside the tensor");
    if (np == 0) return RB_OK;
    RB_REQUIRE(out, "rb_host_ri_mo_pq: out is NULL");
    const i64 plane = (i64)np * nl, cols = (i64)ll * rl;
    RB_REQUIRE(cols == 0 || mo, "rb_host_ri_mo_pq: mo is NULL");
    HOST_CTX(op);
    double *dm = nullptr, *dw = nullptr, *dout;
    RB_TRY(op.alloc(plane * rl, &dm));
    RB_TRY(op.up(dm, mo + plane * r0, plane * rl));
    if (w && cols > 0) {
        RB_TRY(op.alloc(cols, &dw));
        RB_TRY(op.up(dw, w, cols));
    }
    RB_TRY(op.alloc((i64)np * np, &dout));
    RB_TRY(rb_ri_mo_pq(op.ctx, dm, np, np, dm, np, np, nl, rl, l0, ll, 0, rl, dw, 0.0, dout, np));
    RB_TRY(op.down(out, dout, (i64)np * np));
    return op.sync();
}

// ---- eigen-solvers on host buffers (matrix_blas_lapack.rs:319-352, 599-652, 1004-1147, 2123-2185) ----
// _dsyev / lapack_dsyev: dsyev(jobz, 'L', ...) -- the lower triangle of a [n, n] is read; w ascending; z [n, n] (jobz 'V')
extern "C" int rb_host_dsyev(char jobz, int n, const double *a, double *w, double *z)
{
    RB_REQUIRE(n >= 0, "rb_host_dsyev: negative dimension");
    RB_REQUIRE(jobz == 'V' || jobz == 'v' || jobz == 'N' || jobz == 'n', "rb_host_dsyev: jobz must be 'V' or 'N'");
    if (n == 0) return RB_OK;
    const bool want_z = jobz == 'V' || jobz == 'v';
    RB_REQUIRE(a && w && (!want_z || z), "rb_host_dsyev: NULL buffer");
    HOST_CTX(op);
    const i64 nn = (i64)n * n;
    double *da, *dw, *dz = nullptr;
    RB_TRY(op.alloc(nn, &da));
    RB_TRY(op.alloc(n, &dw));
    if (want_z) RB_TRY(op.alloc(nn, &dz));
    RB_TRY(op.up(da, a, nn));
    RB_TRY(rb_dsyev(op.ctx, jobz, 'L', n, da, n, dw, dz, n));
    RB_TRY(op.down(w, dw, n));
    if (want_z) RB_TRY(op.down(z, dz, nn));
    return op.sync();
}

// lapack_dspevx: packed upper input, all eigenpairs; *n_found = n
extern "C" int rb_host_dspevx(int n, const double *ap, double *w, double *z, int *n_found)
{
    RB_REQUIRE(n >= 0, "rb_host_dspevx: negative dimension");
    if (n_found) *n_found = 0;
    if (n == 0) return RB_OK;
    RB_REQUIRE(ap && w && z, "rb_host_dspevx: NULL buffer");
    HOST_CTX(op);
    const i64 nn = (i64)n * n, np = (i64)n * (n + 1) / 2;
    double *dp, *dw, *dz;
    RB_TRY(op.alloc(np, &dp));
    RB_TRY(op.alloc(n, &dw));
    RB_TRY(op.alloc(nn, &dz));
    RB_TRY(op.up(dp, ap, np));
    RB_TRY(rb_dspev(op.ctx, n, dp, dw, dz, n));
    RB_TRY(op.down(w, dw, n));
    RB_TRY(op.down(z, dz, nn));
    if (n_found) *n_found = n;
    return op.sync();
}

// lapack_dspgvx / _dspgvx: A x = lambda B x, packed upper A and B, the num_orb lowest pairs; z [n, num_orb], z^T B z = I
extern "C" int rb_host_dspgvx(int n, const double *ap, const double *bp, int num_orb, double *w, double *z)
{
    RB_REQUIRE(n >= 0 && num_orb >= 0 && num_orb <= n, "rb_host_dspgvx: bad dimensions (n = %d, num_orb = %d)", n, num_orb);
    if (n == 0 || num_orb == 0) return RB_OK;
    RB_REQUIRE(ap && bp && w && z, "rb_host_dspgvx: NULL buffer");
    HOST_CTX(op);
    const i64 np = (i64)n * (n + 1) / 2;
    double *da, *db, *dw, *dz;
    RB_TRY(op.alloc(np, &da));
    RB_TRY(op.alloc(np, &db));
    RB_TRY(op.alloc(num_orb, &dw));
    RB_TRY(op.alloc((i64)n * num_orb, &dz));
    RB_TRY(op.up(da, ap, np));
    RB_TRY(op.up(db, bp, np));
    RB_TRY(rb_dspgv(op.ctx, n, da, db, num_orb, dw, dz, n));
    RB_TRY(op.down(w, dw, num_orb));
    RB_TRY(op.down(z, dz, (i64)n * num_orb));
    return op.sync();
}

// _power / lapack_power: out = sum over eigenvalues >= threshold of lambda^p v v^T
extern "C" int rb_host_power(int n, const double *a, double p, double threshold, double *out, int *n_nonsingular)
{
    RB_REQUIRE(n >= 0, "rb_host_power: negative dimension");
    if (n_nonsingular) *n_nonsingular = 0;
    if (n == 0) return RB_OK;
    RB_REQUIRE(a && out, "rb_host_power: NULL buffer");
    HOST_CTX(op);
    const i64 nn = (i64)n * n;
    double *da, *dout;
    RB_TRY(op.alloc(nn, &da));
    RB_TRY(op.alloc(nn, &dout));
    RB_TRY(op.up(da, a, nn));
    RB_TRY(rb_matrix_power(op.ctx, n, da, n, p, threshold, dout, n, n_nonsingular));
    RB_TRY(op.down(out, dout, nn));
    return op.sync();
}
