// copy_bench.cu -- what limits a 1 read : 1 write SM copy kernel on B200?  (torch / driver D2D memcpy reaches ~6.55 TB/s,
// our double2 grid-stride kernels ~6.05.)  Variants: vector width, loads in flight, cache hints, chunked vs interleaved
// traversal, CTA count, 1-D bulk (TMA) copies.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o copy_bench copy_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <string>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

enum Hint { H_NONE = 0, H_NC_NOALLOC = 1, H_CS = 2, H_EVICT_FIRST = 3 };

template <int HINT>
__device__ __forceinline__ double2 ld16(const double2 *p)
{
    double2 v;
    if (HINT == H_NC_NOALLOC) asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    else if (HINT == H_CS) asm volatile("ld.global.cs.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    else v = *p;
    return v;
}
template <int HINT>
__device__ __forceinline__ void st16(double2 *p, double2 v)
{
    if (HINT == H_CS) asm volatile("st.global.cs.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
    else if (HINT == H_NC_NOALLOC) asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
    else *p = v;
}

// interleaved grid-stride: thread handles U vectors spaced by the grid
template <int U, int HINT>
__global__ void __launch_bounds__(256) k_stride(const double2 *__restrict__ s, double2 *__restrict__ d, size_t n)
{
    size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (U - 1) * stride < n; i += U * stride) {
        double2 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = ld16<HINT>(s + i + u * stride);
#pragma unroll
        for (int u = 0; u < U; ++u) st16<HINT>(d + i + u * stride, v[u]);
    }
    for (; i < n; i += stride) st16<HINT>(d + i, ld16<HINT>(s + i));
}

// chunked: each CTA owns contiguous chunks of CH bytes (taken round-robin); inside, threads stride by blockDim
template <int U, int HINT>
__global__ void __launch_bounds__(256) k_chunk(const double2 *__restrict__ s, double2 *__restrict__ d, size_t n, size_t chunk)
{
    for (size_t c0 = (size_t)blockIdx.x * chunk; c0 < n; c0 += (size_t)gridDim.x * chunk) {
        size_t end = c0 + chunk < n ? c0 + chunk : n;
        size_t i = c0 + threadIdx.x;
        for (; i + (U - 1) * blockDim.x < end; i += U * blockDim.x) {
            double2 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = ld16<HINT>(s + i + u * blockDim.x);
#pragma unroll
            for (int u = 0; u < U; ++u) st16<HINT>(d + i + u * blockDim.x, v[u]);
        }
        for (; i < end; i += blockDim.x) st16<HINT>(d + i, ld16<HINT>(s + i));
    }
}

// 32-byte vectors (sm_100: ld/st.global.v4.f64)
struct alignas(32) d4 { double a, b, c, e; };
template <int U, bool EF>
__global__ void __launch_bounds__(256) k_stride32(const d4 *__restrict__ s, d4 *__restrict__ d, size_t n)
{
    // (Until round 2's last session this loop had no tail: `i + (U - 1) * stride < n` dropped the last, partial pass -- up to 15 % of the
    //  buffer with 32 CTAs per SM -- and every "7.2 - 8.2 TB/s" figure taken from it was inflated by that share.  The last pass is
    //  predicated now, as in the library's rb_copy_flat4_kernel.)
    size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i < n; i += U * stride) {
        d4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            if (i + u * stride >= n) continue;
            unsigned long long a, b, c, e;
            if (EF) asm volatile("ld.global.L1::no_allocate.L2::evict_first.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(e) : "l"(s + i + u * stride));
            else asm volatile("ld.global.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(e) : "l"(s + i + u * stride));
            v[u].a = __longlong_as_double(a); v[u].b = __longlong_as_double(b); v[u].c = __longlong_as_double(c); v[u].e = __longlong_as_double(e);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            if (i + u * stride >= n) continue;
            if (EF) asm volatile("st.global.L1::no_allocate.L2::evict_first.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(d + i + u * stride), "l"(__double_as_longlong(v[u].a)), "l"(__double_as_longlong(v[u].b)), "l"(__double_as_longlong(v[u].c)), "l"(__double_as_longlong(v[u].e)) : "memory");
            else asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(d + i + u * stride), "l"(__double_as_longlong(v[u].a)), "l"(__double_as_longlong(v[u].b)), "l"(__double_as_longlong(v[u].c)), "l"(__double_as_longlong(v[u].e)) : "memory");
        }
    }
}

// traffic-mix probes with 32-byte accesses: write only, 1 read : 2 writes (the unpack pattern), 2 reads : 1 write (axpy), read only
template <int MODE, int U>
__global__ void __launch_bounds__(256) k_mix32(const d4 *__restrict__ s, const d4 *__restrict__ s2, d4 *__restrict__ d, d4 *__restrict__ d2,
                                               size_t n, double *sink)
{
    size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    double acc = 0.0;
    for (; i < n; i += U * stride) { // last pass predicated (see k_stride32)
        d4 v[U], w[U];
#define IN(u) (i + (u) * stride < n)
        if (MODE != 0) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                v[u].a = 0.0; v[u].b = 0.0; v[u].c = 0.0; v[u].e = 0.0;
                if (IN(u)) asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[u].a), "=d"(v[u].b), "=d"(v[u].c), "=d"(v[u].e) : "l"(s + i + u * stride));
            }
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u) { v[u].a = 1.0; v[u].b = 2.0; v[u].c = 3.0; v[u].e = (double)i; }
        }
        if (MODE == 2) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                w[u].a = 0.0; w[u].b = 0.0; w[u].c = 0.0; w[u].e = 0.0;
                if (IN(u)) asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(w[u].a), "=d"(w[u].b), "=d"(w[u].c), "=d"(w[u].e) : "l"(s2 + i + u * stride));
            }
#pragma unroll
            for (int u = 0; u < U; ++u) { v[u].a += w[u].a; v[u].b += w[u].b; v[u].c += w[u].c; v[u].e += w[u].e; }
        }
        if (MODE == 3) {
#pragma unroll
            for (int u = 0; u < U; ++u) acc += v[u].a + v[u].b + v[u].c + v[u].e;
            continue;
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (IN(u)) asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(d + i + u * stride), "d"(v[u].a), "d"(v[u].b), "d"(v[u].c), "d"(v[u].e) : "memory");
        if (MODE == 1) {
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (IN(u)) asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(d2 + i + u * stride), "d"(v[u].a), "d"(v[u].b), "d"(v[u].c), "d"(v[u].e) : "memory");
        }
#undef IN
    }
    if (MODE == 3 && acc == 12345.678) *sink = acc;
}

// 1-D bulk copies through a shared-memory ring, one thread per CTA
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int STAGES>
__global__ void __launch_bounds__(32) k_bulk(const char *__restrict__ s, char *__restrict__ d, size_t bytes, uint32_t piece)
{
    extern __shared__ __align__(128) uint8_t smem[];
    if (threadIdx.x != 0) return;
    const uint32_t base = (smem_u32(smem) + 127u) & ~127u;
    const uint32_t bars = base + STAGES * piece;
    for (int i = 0; i < STAGES; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bars + 8 * i));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const size_t npieces = bytes / piece;
    const size_t first = blockIdx.x, step = gridDim.x;
    const size_t count = first < npieces ? (npieces - first + step - 1) / step : 0;
    auto load = [&](size_t n) {
        const int st = (int)(n % STAGES);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bars + 8 * st), "r"(piece) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(base + st * piece), "l"(s + (first + n * step) * (size_t)piece), "r"(piece), "r"(bars + 8 * st) : "memory");
    };
    for (size_t n = 0; n < STAGES - 1 && n < count; ++n) load(n);
    for (size_t n = 0; n < count; ++n) {
        const int st = (int)(n % STAGES);
        const uint32_t ph = (uint32_t)((n / STAGES) & 1);
        uint32_t done;
        do {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                         : "=r"(done) : "r"(bars + 8 * st), "r"(ph) : "memory");
        } while (!done);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(d + (first + n * step) * (size_t)piece), "r"(base + st * piece), "r"(piece) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (n + STAGES - 1 < count) {
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            load(n + STAGES - 1);
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <typename F>
double time_gbs(F launch, size_t bytes, int reps = 8)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return 2.0 * bytes / (best * 1e-3) / 1e9;
}

int main()
{
    const size_t bytes = (size_t)2 << 30;
    char *s, *d;
    CK(cudaMalloc(&s, bytes)); CK(cudaMalloc(&d, bytes));
    CK(cudaMemset(s, 1, bytes)); CK(cudaMemset(d, 0, bytes));
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    printf("SMs %d, buffer %zu MiB each\n", sms, bytes >> 20);
    printf("%-44s %8.1f GB/s\n", "cudaMemcpy D2D", time_gbs([&] { cudaMemcpyAsync(d, s, bytes, cudaMemcpyDeviceToDevice); }, bytes));
    const size_t n16 = bytes / 16, n32 = bytes / 32;
    for (int mult : {4, 8, 16, 32}) {
        const int grid = sms * mult;
        char name[128];
#define RUN_STRIDE(U, H, HN) snprintf(name, sizeof name, "stride16 U=%d hint=%s grid=%dxSM", U, HN, mult); \
        printf("%-44s %8.1f GB/s\n", name, time_gbs([&] { k_stride<U, H><<<grid, 256>>>((const double2 *)s, (double2 *)d, n16); }, bytes));
        RUN_STRIDE(4, H_NONE, "none") RUN_STRIDE(8, H_NONE, "none") RUN_STRIDE(16, H_NONE, "none")
        RUN_STRIDE(8, H_NC_NOALLOC, "nc.noalloc") RUN_STRIDE(8, H_CS, "cs")
        snprintf(name, sizeof name, "stride32 U=4 grid=%dxSM", mult);
        printf("%-44s %8.1f GB/s\n", name, time_gbs([&] { k_stride32<4, false><<<grid, 256>>>((const d4 *)s, (d4 *)d, n32); }, bytes));
        snprintf(name, sizeof name, "stride32 U=8 grid=%dxSM", mult);
        printf("%-44s %8.1f GB/s\n", name, time_gbs([&] { k_stride32<8, false><<<grid, 256>>>((const d4 *)s, (d4 *)d, n32); }, bytes));
        snprintf(name, sizeof name, "stride32 U=4 evict_first grid=%dxSM", mult);
        printf("%-44s %8.1f GB/s\n", name, time_gbs([&] { k_stride32<4, true><<<grid, 256>>>((const d4 *)s, (d4 *)d, n32); }, bytes));
        snprintf(name, sizeof name, "stride32 U=8 evict_first grid=%dxSM", mult);
        printf("%-44s %8.1f GB/s\n", name, time_gbs([&] { k_stride32<8, true><<<grid, 256>>>((const d4 *)s, (d4 *)d, n32); }, bytes));
        for (size_t chunk_kb : {64, 256, 1024, 4096}) {
            snprintf(name, sizeof name, "chunk16 U=8 none chunk=%zuKB grid=%dxSM", chunk_kb, mult);
            printf("%-44s %8.1f GB/s\n", name, time_gbs([&] { k_chunk<8, H_NONE><<<grid, 256>>>((const double2 *)s, (double2 *)d, n16, chunk_kb * 64); }, bytes));
        }
        snprintf(name, sizeof name, "chunk16 U=8 cs chunk=1024KB grid=%dxSM", mult);
        printf("%-44s %8.1f GB/s\n", name, time_gbs([&] { k_chunk<8, H_CS><<<grid, 256>>>((const double2 *)s, (double2 *)d, n16, 1024 * 64); }, bytes));
    }
    {
        // traffic mixes: bytes counted = all reads + all writes; 1 GiB per stream
        const size_t half = bytes / 2, nh = half / 32;
        const d4 *a = (const d4 *)s, *b = (const d4 *)(s + half);
        d4 *c = (d4 *)d, *e = (d4 *)(d + half);
        double *sink; CK(cudaMalloc(&sink, 8));
        for (int mult : {8, 16, 32}) {
            const int grid = sms * mult;
            char name[128];
            snprintf(name, sizeof name, "mix32 write only U=8 grid=%dxSM", mult);
            printf("%-44s %8.1f GB/s\n", name, 0.5 * time_gbs([&] { k_mix32<0, 8><<<grid, 256>>>(a, b, c, e, nh, sink); }, half));
            snprintf(name, sizeof name, "mix32 read only U=8 grid=%dxSM", mult);
            printf("%-44s %8.1f GB/s\n", name, 0.5 * time_gbs([&] { k_mix32<3, 8><<<grid, 256>>>(a, b, c, e, nh, sink); }, half));
            snprintf(name, sizeof name, "mix32 1r:2w U=4 grid=%dxSM", mult);
            printf("%-44s %8.1f GB/s\n", name, 1.5 * time_gbs([&] { k_mix32<1, 4><<<grid, 256>>>(a, b, c, e, nh, sink); }, half));
            snprintf(name, sizeof name, "mix32 1r:2w U=8 grid=%dxSM", mult);
            printf("%-44s %8.1f GB/s\n", name, 1.5 * time_gbs([&] { k_mix32<1, 8><<<grid, 256>>>(a, b, c, e, nh, sink); }, half));
            snprintf(name, sizeof name, "mix32 2r:1w U=4 grid=%dxSM", mult);
            printf("%-44s %8.1f GB/s\n", name, 1.5 * time_gbs([&] { k_mix32<2, 4><<<grid, 256>>>(a, b, c, e, nh, sink); }, half));
            snprintf(name, sizeof name, "mix32 1r:1w U=8 (1 GiB) grid=%dxSM", mult);
            printf("%-44s %8.1f GB/s\n", name, time_gbs([&] { k_stride32<8, false><<<grid, 256>>>(a, c, nh); }, half));
        }
    }
    for (int per_sm : {1, 2, 4}) {
        for (uint32_t piece : {8192u, 16384u, 32768u}) {
            const int stages = 4;
            if ((size_t)per_sm * (stages * piece + 256) > 220 * 1024) continue;
            char name[128];
            snprintf(name, sizeof name, "bulk1d stages=4 piece=%uKB CTAs/SM=%d", piece >> 10, per_sm);
            CK(cudaFuncSetAttribute(k_bulk<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, stages * piece + 256));
            printf("%-44s %8.1f GB/s\n", name, time_gbs([&] { k_bulk<4><<<sms * per_sm, 32, stages * piece + 256>>>(s, d, bytes, piece); }, bytes));
        }
    }
    for (uint32_t piece : {16384u, 32768u}) {
        char name[128];
        snprintf(name, sizeof name, "bulk1d stages=6 piece=%uKB CTAs/SM=1", piece >> 10);
        CK(cudaFuncSetAttribute(k_bulk<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * piece + 256));
        printf("%-44s %8.1f GB/s\n", name, time_gbs([&] { k_bulk<6><<<sms, 32, 6 * piece + 256>>>(s, d, bytes, piece); }, bytes));
    }
    return 0;
}
