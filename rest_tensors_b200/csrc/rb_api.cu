// rb_api.cu -- context, error reporting, memory helpers, FP64/HBM probes.
#include "rb_common.cuh"
#include <sched.h>
#include <cstring>
#include <cctype>
#include <cstring>

static thread_local char g_err[1024] = "";

void rb_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *rb_last_error(void) { return g_err; }
extern "C" int rb_version(void) { return 100; }

extern "C" int rb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int rb_ctx_create(int device, rb_ctx **out)
{
    RB_REQUIRE(out != nullptr, "rb_ctx_create: out is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        rb_set_error("rb_ctx_create: no CUDA device available (%s); librest_b200 has no CPU fallback",
                     e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        return RB_ERR_CUDA;
    }
    RB_REQUIRE(device >= 0 && device < n, "rb_ctx_create: device %d out of range [0,%d)", device, n);
    RB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    RB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        rb_set_error("rb_ctx_create: device %d is sm_%d%d; this library contains sm_100a code only", device,
                     prop.major, prop.minor);
        return RB_ERR_CUDA;
    }
    rb_ctx *c = new rb_ctx();
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    RB_CUDA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    RB_CUDA(cudaEventCreate(&c->ev0));
    RB_CUDA(cudaEventCreate(&c->ev1));
    RB_CUDA(cudaMalloc((void **)&c->sched, 64 * 2 * sizeof(unsigned long long)));
    RB_CUDA(cudaMemset(c->sched, 0, 64 * 2 * sizeof(unsigned long long)));
    // cuTensorMapEncodeTiled through the runtime's driver entry point lookup (no link against libcuda).
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) c->encode_tiled = (rb_encode_tiled_fn)fn;
    if (const char *lt = getenv("REST_B200_LAYOUT_TMA")) c->layout_path = atoi(lt) != 0 ? 1 : 0;
    else cudaGetLastError();
    *out = c;
    return RB_OK;
}

extern "C" int rb_ctx_destroy(rb_ctx *ctx)
{
    if (!ctx) return RB_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    rb_comm_destroy(ctx);
    rb_eig_cache_free(ctx);
    for (int s = 0; s < 4; ++s) if (ctx->ws[s]) cudaFree(ctx->ws[s]);
    for (void *p : ctx->ws_parked) cudaFree(p);
    if (ctx->sched) cudaFree(ctx->sched);
    for (int i = 0; i < 5; ++i) if (ctx->aux_ev[i]) cudaEventDestroy(ctx->aux_ev[i]);
    if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return RB_OK;
}

extern "C" int rb_ctx_set_stream(rb_ctx *ctx, void *cuda_stream)
{
    RB_REQUIRE(ctx, "rb_ctx_set_stream: ctx is NULL");
    RB_NO_CAPTURE(ctx, "rb_ctx_set_stream");
    ctx->stream = (cudaStream_t)cuda_stream; // NULL is the legacy default stream (what torch uses by default)
    return RB_OK;
}

extern "C" int rb_ctx_use_own_stream(rb_ctx *ctx)
{
    RB_REQUIRE(ctx, "rb_ctx_use_own_stream: ctx is NULL");
    RB_NO_CAPTURE(ctx, "rb_ctx_use_own_stream");
    ctx->stream = ctx->own_stream;
    return RB_OK;
}

extern "C" int rb_ctx_sync(rb_ctx *ctx)
{
    RB_REQUIRE(ctx, "rb_ctx_sync: ctx is NULL");
    RB_NO_CAPTURE(ctx, "rb_ctx_sync");
    RB_CUDA(cudaStreamSynchronize(ctx->stream));
    return RB_OK;
}

extern "C" int rb_ctx_num_sms(rb_ctx *ctx) { return ctx ? ctx->num_sms : 0; }
extern "C" int64_t rb_ctx_launch_count(rb_ctx *ctx) { return ctx ? ctx->launches : 0; }
extern "C" int64_t rb_ctx_tma_layout_count(rb_ctx *ctx) { return ctx ? ctx->tma_layout_launches : 0; }
extern "C" int rb_ctx_set_layout_path(rb_ctx *ctx, int path)
{
    RB_REQUIRE(ctx, "rb_ctx_set_layout_path: ctx is NULL");
    RB_REQUIRE(path == 0 || path == 1, "rb_ctx_set_layout_path: path must be 0 or 1");
    ctx->layout_path = path;
    return RB_OK;
}

extern "C" int rb_ctx_set_gemm_path(rb_ctx *ctx, int path)
{
    RB_REQUIRE(ctx, "rb_ctx_set_gemm_path: ctx is NULL");
    ctx->gemm_path = path;
    return RB_OK;
}

int rb_ws_reserve(rb_ctx *ctx, int slot, i64 bytes, void **out)
{
    if (bytes > ctx->ws_bytes[slot]) {
        if (ctx->capturing) {
            rb_set_error("workspace %d would have to grow (%lld -> %lld bytes) while a graph is being recorded: run the call "
                         "sequence once before rb_graph_begin so that the workspaces have their final size",
                         slot, (long long)ctx->ws_bytes[slot], (long long)bytes);
            return RB_ERR_UNSUPPORTED;
        }
        RB_CUDA(cudaSetDevice(ctx->device));
        RB_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ctx->ws[slot]) {
            if (ctx->live_graphs > 0) ctx->ws_parked.push_back(ctx->ws[slot]); // recordings hold its address: freed with the last of them
            else RB_CUDA(cudaFree(ctx->ws[slot]));
            ctx->ws[slot] = nullptr; ctx->ws_bytes[slot] = 0;
        }
        cudaError_t e = cudaMalloc(&ctx->ws[slot], (size_t)bytes);
        if (e != cudaSuccess) {
            cudaGetLastError();
            rb_set_error("workspace allocation of %lld bytes failed: %s", (long long)bytes, cudaGetErrorString(e));
            ctx->ws_budget = 0; // the cached budget is stale (the application took HBM since): re-query on the next plan
            return RB_ERR_NOMEM;
        }
        ctx->ws_bytes[slot] = bytes;
    }
    *out = ctx->ws[slot];
    return RB_OK;
}

// ---- CUDA-graph capture of a repeated call sequence --------------------------------------------------------------------------
// An SCF iteration issues the same d_P / J / K (and, small systems, ao2mo) calls on the same buffers every time; for small systems the
// cost is the launches, not the kernels (~14 us per GEMM call against < 1 us of DMMA time).  Between
// rb_graph_begin and rb_graph_end the calls on this context are RECORDED (stream capture) instead of run; rb_graph_launch replays
// the recording with one launch.  Everything a call computed on the host is baked in: pointers, shapes, scalars, tensor maps, the
// split-K / stream-K plan, the workspace addresses.  The GEMM's dynamic tile scheduler re-arms itself on the device, so a
// recording can be replayed any number of times.
struct RbGraph {
    cudaGraphExec_t exec;
    i64 kernels; // kernel launches recorded (added to the context's launch count per replay)
};

extern "C" int rb_graph_begin(rb_ctx *ctx)
{
    RB_REQUIRE(ctx, "rb_graph_begin: ctx is NULL");
    RB_REQUIRE(!ctx->capturing, "rb_graph_begin: a recording is already open on this context");
    RB_REQUIRE(ctx->stream != nullptr && ctx->stream != cudaStreamLegacy && ctx->stream != cudaStreamPerThread,
               "rb_graph_begin: the default stream cannot be captured -- rb_ctx_use_own_stream() or rb_ctx_set_stream(a stream you created)");
    RB_CUDA(cudaSetDevice(ctx->device));
    RB_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
    ctx->capturing = 1;
    ctx->capture_launches0 = ctx->launches;
    return RB_OK;
}

extern "C" int rb_graph_end(rb_ctx *ctx, void **graph_out)
{
    RB_REQUIRE(ctx && graph_out, "rb_graph_end: bad arguments");
    RB_REQUIRE(ctx->capturing, "rb_graph_end: no recording is open on this context");
    *graph_out = nullptr;
    ctx->capturing = 0;
    const i64 kernels = ctx->launches - ctx->capture_launches0;
    ctx->launches = ctx->capture_launches0; // recorded, not run
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(ctx->stream, &g);
    if (e != cudaSuccess || !g) {
        (void)cudaGetLastError();
        rb_set_error("rb_graph_end: the recording is invalid (%s): a call inside it needed the host (see its own error)",
                     cudaGetErrorString(e));
        if (g) cudaGraphDestroy(g);
        return RB_ERR_CUDA;
    }
    cudaGraphExec_t ex = nullptr;
    e = cudaGraphInstantiate(&ex, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        rb_set_error("rb_graph_end: cudaGraphInstantiate -> %s", cudaGetErrorString(e));
        return RB_ERR_CUDA;
    }
    (void)cudaGraphUpload(ex, ctx->stream); // first replay then costs what the others do
    (void)cudaGetLastError();
    *graph_out = new RbGraph{ex, kernels};
    ctx->live_graphs++;
    return RB_OK;
}

extern "C" int rb_graph_launch(rb_ctx *ctx, void *graph)
{
    RB_REQUIRE(ctx && graph, "rb_graph_launch: bad arguments");
    RB_REQUIRE(!ctx->capturing, "rb_graph_launch: a recording is open on this context");
    RbGraph *g = (RbGraph *)graph;
    RB_CUDA(cudaSetDevice(ctx->device));
    RB_CUDA(cudaGraphLaunch(g->exec, ctx->stream));
    ctx->launches += g->kernels;
    return RB_OK;
}

extern "C" int64_t rb_graph_kernel_count(void *graph) { return graph ? ((RbGraph *)graph)->kernels : 0; }

extern "C" int rb_graph_free(rb_ctx *ctx, void *graph)
{
    RB_REQUIRE(ctx, "rb_graph_free: ctx is NULL");
    if (!graph) return RB_OK;
    RbGraph *g = (RbGraph *)graph;
    RB_CUDA(cudaSetDevice(ctx->device));
    RB_CUDA(cudaStreamSynchronize(ctx->stream)); // a replay may still be running
    cudaGraphExecDestroy(g->exec);
    delete g;
    if (--ctx->live_graphs <= 0) {
        ctx->live_graphs = 0;
        for (void *p : ctx->ws_parked) cudaFree(p);
        ctx->ws_parked.clear();
    }
    return RB_OK;
}

// Test hook: fill every workspace with the all-ones bit pattern (a NaN as a double, -1 as an integer) on the context's stream.  A kernel
// that reads workspace memory it (or an earlier kernel of the same call) did not write then shows up as NaN in the result instead of
// passing on whatever the previous call left there (tests/test_gpu_contractions.py::test_partials_never_read_stale_workspace).
extern "C" int rb_ctx_poison_workspaces(rb_ctx *ctx)
{
    RB_REQUIRE(ctx, "rb_ctx_poison_workspaces: ctx is NULL");
    RB_CUDA(cudaSetDevice(ctx->device));
    for (int s = 0; s < 4; ++s)
        if (ctx->ws[s] && ctx->ws_bytes[s] > 0) RB_CUDA(cudaMemsetAsync(ctx->ws[s], 0xff, (size_t)ctx->ws_bytes[s], ctx->stream));
    return RB_OK;
}

extern "C" int rb_dev_alloc(rb_ctx *ctx, int64_t bytes, void **out)
{
    RB_REQUIRE(ctx && out && bytes >= 0, "rb_dev_alloc: bad arguments");
    RB_CUDA(cudaSetDevice(ctx->device));
    *out = nullptr;
    if (bytes == 0) return RB_OK;
    cudaError_t e = cudaMalloc(out, (size_t)bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        rb_set_error("rb_dev_alloc(%lld) failed: %s", (long long)bytes, cudaGetErrorString(e));
        return RB_ERR_NOMEM;
    }
    return RB_OK;
}
extern "C" int rb_dev_free(rb_ctx *ctx, void *p)
{
    RB_REQUIRE(ctx, "rb_dev_free: ctx is NULL");
    if (p) RB_CUDA(cudaFree(p));
    return RB_OK;
}
// ---- peer memory (NVLink / NVSwitch) -----------------------------------------------------------------------------
// Kernels of this library take plain global pointers, and TMA tensor maps are encoded over plain global addresses, so
// an operand may live in another GPU's HBM once the mapping exists: same process -> rb_peer_enable; another process
// (one rank per GPU) -> export the cudaMalloc'd block as a 64-byte handle and open it on the reading rank.
extern "C" int rb_peer_enable(rb_ctx *ctx, int peer_device)
{
    RB_REQUIRE(ctx, "rb_peer_enable: ctx is NULL");
    if (peer_device == ctx->device) return RB_OK;
    RB_CUDA(cudaSetDevice(ctx->device));
    int can = 0;
    RB_CUDA(cudaDeviceCanAccessPeer(&can, ctx->device, peer_device));
    if (!can) {
        rb_set_error("rb_peer_enable: device %d cannot access device %d", ctx->device, peer_device);
        return RB_ERR_UNSUPPORTED;
    }
    cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return RB_OK; }
    RB_CUDA(e);
    return RB_OK;
}

extern "C" int rb_ipc_export(rb_ctx *ctx, void *dev_ptr, unsigned char handle[64], int64_t *offset_out)
{
    RB_REQUIRE(ctx && dev_ptr && handle, "rb_ipc_export: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    RB_CUDA(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    RB_CUDA(cudaIpcGetMemHandle(&h, dev_ptr)); // the handle names the whole allocation that contains dev_ptr
    memcpy(handle, &h, 64);
    if (offset_out) {
        // offset of dev_ptr inside that allocation (a sub-block of a caching allocator's segment is exported this way)
        typedef CUresult (*range_fn)(CUdeviceptr *, size_t *, CUdeviceptr);
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
            cudaGetLastError();
            rb_set_error("rb_ipc_export: cuMemGetAddressRange is not available");
            return RB_ERR_CUDA;
        }
        CUdeviceptr base = 0;
        size_t size = 0;
        const CUresult r = ((range_fn)fn)(&base, &size, (CUdeviceptr)(uintptr_t)dev_ptr);
        if (r != CUDA_SUCCESS) {
            rb_set_error("rb_ipc_export: cuMemGetAddressRange failed (%d)", (int)r);
            return RB_ERR_CUDA;
        }
        *offset_out = (int64_t)((uintptr_t)dev_ptr - (uintptr_t)base);
    }
    return RB_OK;
}

extern "C" int rb_ipc_open(rb_ctx *ctx, const unsigned char handle[64], void **out)
{
    RB_REQUIRE(ctx && handle && out, "rb_ipc_open: bad arguments");
    RB_CUDA(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    *out = nullptr;
    RB_CUDA(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
    return RB_OK;
}

extern "C" int rb_ipc_close(rb_ctx *ctx, void *ptr)
{
    RB_REQUIRE(ctx, "rb_ipc_close: ctx is NULL");
    if (!ptr) return RB_OK;
    RB_CUDA(cudaSetDevice(ctx->device));
    RB_CUDA(cudaIpcCloseMemHandle(ptr));
    return RB_OK;
}

extern "C" int rb_host_alloc_pinned(int64_t bytes, void **out)
{
    RB_REQUIRE(out && bytes >= 0, "rb_host_alloc_pinned: bad arguments");
    *out = nullptr;
    if (bytes == 0) return RB_OK;
    RB_CUDA(cudaMallocHost(out, (size_t)bytes));
    return RB_OK;
}
extern "C" int rb_host_free_pinned(void *p)
{
    if (p) RB_CUDA(cudaFreeHost(p));
    return RB_OK;
}
// Page-lock a buffer the CALLER owns (e.g. the Vec<f64> that holds ri3ao for the whole SCF run) so that the host-pointer entry
// points stream it at the pinned rate instead of bouncing it; the caller must unregister it before freeing it.
extern "C" int rb_host_register(void *p, int64_t bytes)
{
    RB_REQUIRE(p && bytes > 0, "rb_host_register: bad arguments");
    RB_CUDA(cudaHostRegister(p, (size_t)bytes, cudaHostRegisterPortable));
    return RB_OK;
}
extern "C" int rb_host_unregister(void *p)
{
    RB_REQUIRE(p, "rb_host_unregister: NULL pointer");
    RB_CUDA(cudaHostUnregister(p));
    return RB_OK;
}
extern "C" int rb_memcpy_h2d(rb_ctx *ctx, void *dst, const void *src, int64_t bytes)
{
    RB_REQUIRE(ctx, "rb_memcpy_h2d: ctx is NULL");
    if (bytes > 0) RB_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyHostToDevice, ctx->stream));
    return RB_OK;
}
extern "C" int rb_memcpy_d2h(rb_ctx *ctx, void *dst, const void *src, int64_t bytes)
{
    RB_REQUIRE(ctx, "rb_memcpy_d2h: ctx is NULL");
    if (bytes > 0) RB_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return RB_OK;
}

// ---- default context for host-pointer entry points -----------------------------------------------------
static rb_ctx *g_default = nullptr;
static std::mutex g_default_mutex;
std::mutex &rb_default_mutex(void) { return g_default_mutex; }
rb_ctx *rb_default_ctx(void)
{
    if (!g_default) {
        int dev = 0;
        const char *env = getenv("REST_B200_DEVICE");
        if (env) dev = atoi(env);
        rb_ctx *c = nullptr;
        if (rb_ctx_create(dev, &c) != RB_OK) return nullptr;
        g_default = c;
    }
    cudaSetDevice(g_default->device);
    return g_default;
}

// ---- host placement -----------------------------------------------------------------------------------------
// sysfs: /sys/bus/pci/devices/<domain:bus:dev.fn>/{numa_node,local_cpulist}
static bool read_line(const char *path, char *buf, size_t n)
{
    FILE *f = fopen(path, "r");
    if (!f) return false;
    bool ok = fgets(buf, (int)n, f) != nullptr;
    fclose(f);
    return ok;
}

extern "C" int rb_bind_host_to_device_numa(int device, int *node_out)
{
    if (node_out) *node_out = -1;
    char bus[32] = {0};
    RB_CUDA(cudaDeviceGetPCIBusId(bus, sizeof bus, device));
    for (char *c = bus; *c; ++c) *c = (char)tolower((unsigned char)*c);
    char path[128], line[4096];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
    int node = -1;
    if (read_line(path, line, sizeof line)) node = atoi(line);
    if (node_out) *node_out = node;
    if (node < 0) return RB_OK; // no NUMA information: leave the affinity alone
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/local_cpulist", bus);
    if (!read_line(path, line, sizeof line)) return RB_OK;
    cpu_set_t want, have, both;
    CPU_ZERO(&want);
    for (char *tok = strtok(line, ",\n"); tok; tok = strtok(nullptr, ",\n")) { // "0-31,64-95"
        int lo = 0, hi = 0;
        if (sscanf(tok, "%d-%d", &lo, &hi) == 2) { for (int c = lo; c <= hi && c < CPU_SETSIZE; ++c) CPU_SET(c, &want); }
        else if (sscanf(tok, "%d", &lo) == 1 && lo < CPU_SETSIZE) CPU_SET(lo, &want);
    }
    if (sched_getaffinity(0, sizeof have, &have) != 0) return RB_OK;
    CPU_AND(&both, &want, &have); // never widen what the container / launcher allowed
    if (CPU_COUNT(&both) == 0) return RB_OK;
    if (sched_setaffinity(0, sizeof both, &both) != 0) return RB_OK;
    return RB_OK;
}

// ---- probes ------------------------------------------------------------------------------------------------
// Register-resident FP64 pipe probes: the roofline denominator for the DMMA kernels is measured, not assumed.
template <int KIND>
__global__ void __launch_bounds__(256) rb_fp64_probe_kernel(double *sink, int iters, double seed)
{
    // 16 independent accumulator tiles per warp keep the pipe full regardless of its latency.
    double c[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) { c[i][0] = seed * i; c[i][1] = seed; }
    double a = seed + threadIdx.x * 1e-9, b = seed - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (KIND == 0) {
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][0]), "+d"(c[i][1])
                             : "d"(a), "d"(b));
            } else {
                c[i][0] = fma(a, b, c[i][0]);
                c[i][1] = fma(b, a, c[i][1]);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) sink[0] = s; // never true; keeps the loop alive
}

extern "C" int rb_fp64_peak_probe(rb_ctx *ctx, int kind, int iters, double *tflops_out, double *ms_out)
{
    RB_REQUIRE(ctx && tflops_out, "rb_fp64_peak_probe: bad arguments");
    RB_NO_CAPTURE(ctx, "rb_fp64_peak_probe");
    RB_REQUIRE(kind == 0 || kind == 1, "rb_fp64_peak_probe: kind must be 0 (DMMA) or 1 (DFMA)");
    RB_CUDA(cudaSetDevice(ctx->device));
    void *sink;
    RB_TRY(rb_ws_reserve(ctx, 3, 256, &sink));
    const int blocks = ctx->num_sms * 2, threads = 256;
    for (int rep = 0; rep < 2; ++rep) { // rep 0 = warm-up
        RB_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
        if (kind == 0) rb_fp64_probe_kernel<0><<<blocks, threads, 0, ctx->stream>>>((double *)sink, iters, 1e-3);
        else rb_fp64_probe_kernel<1><<<blocks, threads, 0, ctx->stream>>>((double *)sink, iters, 1e-3);
        RB_LAUNCHED(ctx);
        RB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
        RB_CUDA(cudaEventSynchronize(ctx->ev1));
    }
    float ms = 0;
    RB_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    // DMMA m8n8k4: 8*8*4 FMA = 512 flop per warp-instruction; DFMA: 2 flop per thread-instruction, 2 per tile here.
    double flop;
    if (kind == 0) flop = (double)blocks * (threads / 32) * (double)iters * 16.0 * 512.0;
    else flop = (double)blocks * threads * (double)iters * 16.0 * 2.0 * 2.0;
    *tflops_out = flop / (ms * 1e-3) / 1e12;
    if (ms_out) *ms_out = ms;
    return RB_OK;
}

__global__ void __launch_bounds__(256) rb_copy_probe_kernel(const double2 *__restrict__ src, double2 *__restrict__ dst,
                                                            i64 n2)
{
    const i64 stride = (i64)gridDim.x * blockDim.x;
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n2; i += 4 * stride) { // 4 independent 16-byte loads in flight per thread
        const double2 v0 = src[i], v1 = src[i + stride], v2 = src[i + 2 * stride], v3 = src[i + 3 * stride];
        dst[i] = v0; dst[i + stride] = v1; dst[i + 2 * stride] = v2; dst[i + 3 * stride] = v3;
    }
    for (; i < n2; i += stride) dst[i] = src[i];
}

extern "C" int rb_hbm_copy_probe(rb_ctx *ctx, int64_t bytes, int iters, double *gbs_out)
{
    RB_REQUIRE(ctx && gbs_out && bytes >= 32 && iters > 0, "rb_hbm_copy_probe: bad arguments");
    RB_NO_CAPTURE(ctx, "rb_hbm_copy_probe");
    RB_CUDA(cudaSetDevice(ctx->device));
    void *buf;
    i64 half = (bytes / 2) & ~(i64)15;
    RB_TRY(rb_ws_reserve(ctx, 3, half * 2, &buf));
    const double2 *src = (const double2 *)buf;
    double2 *dst = (double2 *)((char *)buf + half);
    i64 n2 = half / 16;
    int blocks = ctx->num_sms * 8;
    rb_copy_probe_kernel<<<blocks, 256, 0, ctx->stream>>>(src, dst, n2);
    RB_LAUNCHED(ctx);
    RB_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    for (int i = 0; i < iters; ++i) {
        rb_copy_probe_kernel<<<blocks, 256, 0, ctx->stream>>>(src, dst, n2);
        RB_LAUNCHED(ctx);
    }
    RB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    RB_CUDA(cudaEventSynchronize(ctx->ev1));
    float ms = 0;
    RB_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    *gbs_out = (double)half * 2.0 * iters / (ms * 1e-3) / 1e9;
    return RB_OK;
}


// ---- PCIe probe: what the host-pointer paths can hope for on this box -------------------------------------------------
// mode 0: H2D contiguous (pinned)      1: D2H contiguous (pinned)        2: D2H cudaMemcpy2D with rows of `width` bytes
// mode 3: H2D and D2H contiguous at the same time (full duplex)          4: D2H by kernel stores into mapped pinned memory
// mode 5: like 4 but rows of `width` bytes scattered at twice the pitch (the P-chunk pattern of ri3mo)
__global__ void __launch_bounds__(256) rb_zero_copy_store_kernel(const double2 *__restrict__ src, double2 *__restrict__ dst,
                                                                 i64 n2, i64 row2, i64 pitch2)
{
    i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
        i64 r = i / row2, c = i - r * row2;
        dst[r * pitch2 + c] = src[i];
    }
}

extern "C" int rb_pcie_probe(rb_ctx *ctx, int mode, int64_t bytes, int64_t width, int iters, double *gbs_out)
{
    RB_REQUIRE(ctx && gbs_out && bytes >= 4096 && iters > 0 && mode >= 0 && mode <= 5, "rb_pcie_probe: bad arguments");
    RB_NO_CAPTURE(ctx, "rb_pcie_probe");
    RB_CUDA(cudaSetDevice(ctx->device));
    if (width <= 0) width = 2048;
    bytes = (bytes / width) * width;
    void *h = nullptr, *h2 = nullptr, *d = nullptr, *d2 = nullptr;
    const i64 hbytes = (mode == 5 || mode == 2) ? bytes * 2 : bytes;
    RB_CUDA(cudaMallocHost(&h, (size_t)hbytes));
    RB_CUDA(cudaMalloc(&d, (size_t)bytes));
    if (mode == 3) { RB_CUDA(cudaMallocHost(&h2, (size_t)bytes)); RB_CUDA(cudaMalloc(&d2, (size_t)bytes)); }
    cudaStream_t s2 = nullptr;
    RB_CUDA(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    RB_CUDA(cudaMemsetAsync(d, 0, (size_t)bytes, ctx->stream));
    int status = RB_OK;
    for (int rep = 0; rep < 2 && status == RB_OK; ++rep) { // rep 0 = warm-up
        RB_CUDA(cudaStreamSynchronize(ctx->stream));
        RB_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
        for (int it = 0; it < (rep == 0 ? 1 : iters); ++it) {
            switch (mode) {
            case 0: RB_CUDA(cudaMemcpyAsync(d, h, (size_t)bytes, cudaMemcpyHostToDevice, ctx->stream)); break;
            case 1: RB_CUDA(cudaMemcpyAsync(h, d, (size_t)bytes, cudaMemcpyDeviceToHost, ctx->stream)); break;
            case 2: RB_CUDA(cudaMemcpy2DAsync(h, (size_t)width * 2, d, (size_t)width, (size_t)width, (size_t)(bytes / width),
                                              cudaMemcpyDeviceToHost, ctx->stream)); break;
            case 3:
                RB_CUDA(cudaMemcpyAsync(d, h, (size_t)bytes, cudaMemcpyHostToDevice, ctx->stream));
                RB_CUDA(cudaMemcpyAsync(h2, d2, (size_t)bytes, cudaMemcpyDeviceToHost, s2));
                break;
            default: {
                i64 row2 = (mode == 5) ? width / 16 : bytes / 16, pitch2 = (mode == 5) ? width / 8 : bytes / 16;
                rb_zero_copy_store_kernel<<<64, 256, 0, ctx->stream>>>((const double2 *)d, (double2 *)h, bytes / 16, row2, pitch2);
                RB_LAUNCHED(ctx);
            }
            }
        }
        if (mode == 3) RB_CUDA(cudaStreamSynchronize(s2));
        RB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
        RB_CUDA(cudaEventSynchronize(ctx->ev1));
    }
    float ms = 0;
    RB_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    *gbs_out = (double)bytes * iters * (mode == 3 ? 2.0 : 1.0) / (ms * 1e-3) / 1e9;
    cudaStreamDestroy(s2);
    cudaFreeHost(h); cudaFree(d);
    if (h2) cudaFreeHost(h2);
    if (d2) cudaFree(d2);
    return status;
}
