// rb_comm.cu -- the collectives of the P-sharded RI path behind the C ABI.
//
// rank r owns the slabs iter_auxbas(P_lo..P_hi) (reference src/ri.rs:190-198); ao2mo and d_P need no exchange,
// J and K are per-rank partial sums completed by ONE all-reduce(sum, f64) each, and the full d vector is an all-gather of
// <= 38 KB.  The reference has no communication layer at all, so a Rust host that adopts this library needs these three
// collectives from the library itself: they are NCCL calls (NVLink 5 / NVSwitch; NVLS in-switch reduction when NCCL
// selects it) issued on the context's stream, so they order with the kernels around them without host synchronisation.
//
// NCCL is bound at run time (dlopen), not at link time: the library keeps loading on hosts without NCCL (single-GPU
// use), and inside a process that already carries an NCCL (PyTorch bundles its own libnccl.so.2) the SAME copy is reused
// instead of a second one being mapped next to it.  Search order: $REST_B200_NCCL_LIB, an already loaded libnccl.so.2,
// then the dynamic loader's search path.  Only the few prototypes used are declared here (nccl.h 2.27: the unique id is a
// 128-byte struct passed by value; ncclSum = 0, ncclFloat64 = 8).
#include "rb_common.cuh"
#include <dlfcn.h>
#include <cstring>

namespace {

struct NcclId { char internal[128]; };
typedef void *NcclComm;
typedef int (*fn_get_version)(int *);
typedef int (*fn_get_unique_id)(NcclId *);
typedef int (*fn_comm_init_rank)(NcclComm *, int, NcclId, int);
typedef int (*fn_comm_init_all)(NcclComm *, int, const int *);
typedef int (*fn_comm_destroy)(NcclComm);
typedef const char *(*fn_get_error_string)(int);
typedef int (*fn_all_reduce)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t);
typedef int (*fn_broadcast)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t);
typedef int (*fn_group)(void);

struct NcclApi {
    void *handle = nullptr;
    fn_get_version get_version = nullptr;
    fn_get_unique_id get_unique_id = nullptr;
    fn_comm_init_rank comm_init_rank = nullptr;
    fn_comm_init_all comm_init_all = nullptr;
    fn_comm_destroy comm_destroy = nullptr;
    fn_get_error_string get_error_string = nullptr;
    fn_all_reduce all_reduce = nullptr;
    fn_broadcast broadcast = nullptr;
    fn_group group_start = nullptr, group_end = nullptr;
    char path[256] = "";
};

NcclApi g_nccl;
std::mutex g_nccl_mutex;
constexpr int NCCL_SUM = 0, NCCL_F64 = 8;

int nccl_load(void)
{
    std::lock_guard<std::mutex> lock(g_nccl_mutex);
    if (g_nccl.handle) return RB_OK;
    void *h = nullptr;
    const char *env = getenv("REST_B200_NCCL_LIB");
    const char *tried = "libnccl.so.2";
    if (env && *env) { h = dlopen(env, RTLD_NOW | RTLD_GLOBAL); tried = env; }
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        rb_set_error("rb_comm: cannot load NCCL (%s): %s; set REST_B200_NCCL_LIB to libnccl.so.2", tried, dlerror());
        return RB_ERR_UNSUPPORTED;
    }
    NcclApi a;
    a.handle = h;
#define RB_SYM(field, type, name)                                                                                       \
    a.field = (type)dlsym(h, name);                                                                                     \
    if (!a.field) { rb_set_error("rb_comm: %s is missing from the NCCL library", name); return RB_ERR_UNSUPPORTED; }
    RB_SYM(get_version, fn_get_version, "ncclGetVersion")
    RB_SYM(get_unique_id, fn_get_unique_id, "ncclGetUniqueId")
    RB_SYM(comm_init_rank, fn_comm_init_rank, "ncclCommInitRank")
    RB_SYM(comm_init_all, fn_comm_init_all, "ncclCommInitAll")
    RB_SYM(comm_destroy, fn_comm_destroy, "ncclCommDestroy")
    RB_SYM(get_error_string, fn_get_error_string, "ncclGetErrorString")
    RB_SYM(all_reduce, fn_all_reduce, "ncclAllReduce")
    RB_SYM(broadcast, fn_broadcast, "ncclBroadcast")
    RB_SYM(group_start, fn_group, "ncclGroupStart")
    RB_SYM(group_end, fn_group, "ncclGroupEnd")
#undef RB_SYM
    Dl_info info;
    if (dladdr((void *)a.all_reduce, &info) && info.dli_fname) snprintf(a.path, sizeof a.path, "%s", info.dli_fname);
    g_nccl = a;
    return RB_OK;
}

#define RB_NCCL(call)                                                                                                   \
    do {                                                                                                                \
        const int _r = (call);                                                                                          \
        if (_r != 0) {                                                                                                  \
            rb_set_error("%s:%d: %s -> NCCL error %d (%s)", __FILE__, __LINE__, #call, _r, g_nccl.get_error_string(_r)); \
            return RB_ERR_CUDA;                                                                                         \
        }                                                                                                               \
    } while (0)

} // namespace

extern "C" int rb_comm_nccl_version(int *version_out, char *path_out, int path_len)
{
    RB_REQUIRE(version_out, "rb_comm_nccl_version: version_out is NULL");
    RB_TRY(nccl_load());
    RB_NCCL(g_nccl.get_version(version_out));
    if (path_out && path_len > 0) snprintf(path_out, (size_t)path_len, "%s", g_nccl.path);
    return RB_OK;
}

extern "C" int rb_comm_unique_id(unsigned char id[128])
{
    RB_REQUIRE(id, "rb_comm_unique_id: id is NULL");
    RB_TRY(nccl_load());
    NcclId u;
    RB_NCCL(g_nccl.get_unique_id(&u));
    memcpy(id, u.internal, 128);
    return RB_OK;
}

extern "C" int rb_comm_init_rank(rb_ctx *ctx, int rank, int world, const unsigned char id[128])
{
    RB_REQUIRE(ctx && id, "rb_comm_init_rank: NULL argument");
    RB_NO_CAPTURE(ctx, "rb_comm_init_rank");
    RB_REQUIRE(world >= 1 && rank >= 0 && rank < world, "rb_comm_init_rank: rank %d outside world %d", rank, world);
    RB_REQUIRE(!ctx->comm, "rb_comm_init_rank: the context already has a communicator");
    RB_TRY(nccl_load());
    RB_CUDA(cudaSetDevice(ctx->device));
    NcclId u;
    memcpy(u.internal, id, 128);
    NcclComm c = nullptr;
    RB_NCCL(g_nccl.comm_init_rank(&c, world, u, rank));
    ctx->comm = c; ctx->comm_rank = rank; ctx->comm_world = world;
    return RB_OK;
}

// One host process driving n devices (a Rust host with one context per GPU): ncclCommInitAll over the contexts' devices.
extern "C" int rb_comm_init_all(rb_ctx *const *ctxs, int n)
{
    RB_REQUIRE(ctxs && n >= 1 && n <= 64, "rb_comm_init_all: bad arguments");
    int devs[64];
    for (int i = 0; i < n; ++i) {
        RB_REQUIRE(ctxs[i], "rb_comm_init_all: context %d is NULL", i);
        RB_REQUIRE(!ctxs[i]->comm, "rb_comm_init_all: context %d already has a communicator", i);
        devs[i] = ctxs[i]->device;
        for (int j = 0; j < i; ++j) RB_REQUIRE(devs[j] != devs[i], "rb_comm_init_all: device %d listed twice", devs[i]);
    }
    RB_TRY(nccl_load());
    NcclComm comms[64];
    RB_NCCL(g_nccl.comm_init_all(comms, n, devs));
    for (int i = 0; i < n; ++i) { ctxs[i]->comm = comms[i]; ctxs[i]->comm_rank = i; ctxs[i]->comm_world = n; }
    return RB_OK;
}

extern "C" int rb_comm_destroy(rb_ctx *ctx)
{
    RB_REQUIRE(ctx, "rb_comm_destroy: ctx is NULL");
    if (!ctx->comm) return RB_OK;
    RB_CUDA(cudaSetDevice(ctx->device));
    RB_CUDA(cudaStreamSynchronize(ctx->stream));
    NcclComm c = ctx->comm;
    ctx->comm = nullptr; ctx->comm_rank = 0; ctx->comm_world = 1;
    RB_NCCL(g_nccl.comm_destroy(c));
    return RB_OK;
}

extern "C" int rb_comm_rank(rb_ctx *ctx) { return ctx ? ctx->comm_rank : 0; }
extern "C" int rb_comm_world(rb_ctx *ctx) { return ctx ? ctx->comm_world : 1; }

// Single-process multi-device callers bracket the per-context collective calls of one step (NCCL group semantics).
extern "C" int rb_comm_group_start(void)
{
    RB_TRY(nccl_load());
    RB_NCCL(g_nccl.group_start());
    return RB_OK;
}
extern "C" int rb_comm_group_end(void)
{
    RB_TRY(nccl_load());
    RB_NCCL(g_nccl.group_end());
    return RB_OK;
}

// In-place sum over the ranks of the context's communicator, asynchronous on the context's stream.  A context without a
// communicator is a world of one: nothing to do.  Recordable (rb_graph_begin): NCCL captures its collectives into the graph; every
// rank must then record and replay the same sequence.
extern "C" int rb_allreduce_sum(rb_ctx *ctx, double *buf, int64_t n)
{
    RB_REQUIRE(ctx, "rb_allreduce_sum: ctx is NULL");
    RB_REQUIRE(n >= 0, "rb_allreduce_sum: negative length");
    if (!ctx->comm || ctx->comm_world == 1 || n == 0) return RB_OK;
    RB_REQUIRE(buf, "rb_allreduce_sum: buf is NULL");
    RB_CUDA(cudaSetDevice(ctx->device));
    RB_NCCL(g_nccl.all_reduce(buf, buf, (size_t)n, NCCL_F64, NCCL_SUM, (NcclComm)ctx->comm, ctx->stream));
    return RB_OK;
}

// full[0..naux) on every rank from the per-rank pieces d_local[P_lo..P_hi) of the shard_range partition (pieces differ in
// length by at most one, so this is `world` broadcasts in one group rather than a fixed-count all-gather).
extern "C" int rb_allgather_shards(rb_ctx *ctx, const double *local, double *full, int64_t naux)
{
    RB_REQUIRE(ctx && naux >= 0, "rb_allgather_shards: bad arguments");
    if (naux == 0) return RB_OK;
    RB_REQUIRE(full, "rb_allgather_shards: full is NULL");
    RB_CUDA(cudaSetDevice(ctx->device));
    const i64 world = ctx->comm ? ctx->comm_world : 1, rank = ctx->comm ? ctx->comm_rank : 0;
    const i64 lo = rank * naux / world, hi = (rank + 1) * naux / world;
    if (hi > lo) {
        RB_REQUIRE(local, "rb_allgather_shards: local is NULL");
        if (local != full + lo)
            RB_CUDA(cudaMemcpyAsync(full + lo, local, (size_t)(hi - lo) * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    if (world == 1) return RB_OK;
    RB_NCCL(g_nccl.group_start());
    for (i64 s = 0; s < world; ++s) {
        const i64 slo = s * naux / world, shi = (s + 1) * naux / world;
        if (shi > slo) {
            const int r = g_nccl.broadcast(full + slo, full + slo, (size_t)(shi - slo), NCCL_F64, (int)s, (NcclComm)ctx->comm, ctx->stream);
            if (r != 0) {
                g_nccl.group_end();
                rb_set_error("rb_allgather_shards: ncclBroadcast -> NCCL error %d (%s)", r, g_nccl.get_error_string(r));
                return RB_ERR_CUDA;
            }
        }
    }
    RB_NCCL(g_nccl.group_end());
    return RB_OK;
}

// J and K of the P-sharded tensor, complete on every rank: the local partial sum followed by the all-reduce on the same
// stream (the only collectives on the hot path).
extern "C" int rb_ri_j_allreduce(rb_ctx *ctx, const double *ri3ao, const double *d, double *j, int nb, int nx)
{
    RB_TRY(rb_ri_j(ctx, ri3ao, d, j, nb, nx));
    return rb_allreduce_sum(ctx, j, (int64_t)nb * nb);
}

extern "C" int rb_ri_k_allreduce(rb_ctx *ctx, const double *ri3ao, const double *ct, int no, double *k, int nb, int nx)
{
    RB_TRY(rb_ri_k(ctx, ri3ao, ct, no, k, nb, nx));
    return rb_allreduce_sum(ctx, k, (int64_t)nb * nb);
}
