// size_sweep.cu -- does the 32-byte grid-stride copy (7.6 TB/s on 2 GiB buffers) keep its rate on the 128 MB .. 1 GiB buffers the library's
// callers use, or is the difference a fixed per-launch cost?  Times ONE launch per event pair (best of 9) and a back-to-back stream of launches
// over distinct buffer pairs inside one event pair, for several buffer sizes; also cudaMemcpyAsync D2D for the same sizes.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o size_sweep size_sweep.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
struct alignas(32) d4 { double a, b, c, e; };
__device__ __forceinline__ d4 ld32(const d4 *p)
{
    d4 v;
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.a), "=d"(v.b), "=d"(v.c), "=d"(v.e) : "l"(p));
    return v;
}
__device__ __forceinline__ void st32(d4 *p, const d4 &v)
{
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.a), "d"(v.b), "d"(v.c), "d"(v.e) : "memory");
}
// the library's rb_copy_flat4_kernel: predicated 8-deep grid-stride passes
__global__ void __launch_bounds__(256) k_flat(const d4 *__restrict__ s, d4 *__restrict__ d, long long n)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += 8 * stride) {
        d4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) if (i + u * stride < n) v[u] = ld32(s + i + u * stride);
#pragma unroll
        for (int u = 0; u < 8; ++u) if (i + u * stride < n) st32(d + i + u * stride, v[u]);
    }
}
int main()
{
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const size_t total = (size_t)4 << 30; // one 4 GiB arena for sources, one for destinations
    char *S, *D;
    CK(cudaMalloc(&S, total)); CK(cudaMalloc(&D, total));
    CK(cudaMemset(S, 1, total)); CK(cudaMemset(D, 0, total));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    printf("%10s %6s | %12s %12s | %12s %12s | %12s\n", "MiB", "grid", "single us", "single GB/s", "stream us", "stream GB/s", "memcpy GB/s");
    for (size_t mib : {64, 128, 256, 512, 1024, 2048}) {
        const size_t bytes = mib << 20;
        const long long n = (long long)(bytes / 32);
        const int sets = (int)(total / bytes);
        for (int mult : {8, 16, 32}) {
            long long blocks = (n + 2047) / 2048;
            if (blocks > (long long)sms * mult) blocks = (long long)sms * mult;
            k_flat<<<(unsigned)blocks, 256>>>((const d4 *)S, (d4 *)D, n); k_flat<<<(unsigned)blocks, 256>>>((const d4 *)S, (d4 *)D, n);
            CK(cudaDeviceSynchronize());
            float best = 1e30f;
            for (int r = 0; r < 9; ++r) {
                const size_t off = (size_t)(r % sets) * bytes;
                CK(cudaEventRecord(e0)); k_flat<<<(unsigned)blocks, 256>>>((const d4 *)(S + off), (d4 *)(D + off), n); CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
            }
            // stream: >= 4 GiB of traffic per direction back to back
            const int calls = sets < 4 ? 4 : sets;
            float sbest = 1e30f;
            for (int rep = 0; rep < 3; ++rep) {
                CK(cudaEventRecord(e0));
                for (int c = 0; c < calls; ++c) { const size_t off = (size_t)(c % sets) * bytes; k_flat<<<(unsigned)blocks, 256>>>((const d4 *)(S + off), (d4 *)(D + off), n); }
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms / calls < sbest) sbest = ms / calls;
            }
            float mbest = 1e30f;
            for (int r = 0; r < 5; ++r) {
                const size_t off = (size_t)(r % sets) * bytes;
                CK(cudaEventRecord(e0)); CK(cudaMemcpyAsync(D + off, S + off, bytes, cudaMemcpyDeviceToDevice)); CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < mbest) mbest = ms;
            }
            printf("%10zu %4dxSM | %12.1f %12.1f | %12.1f %12.1f | %12.1f\n", mib, mult, best * 1e3, 2.0 * bytes / (best * 1e-3) / 1e9, sbest * 1e3,
                   2.0 * bytes / (sbest * 1e-3) / 1e9, 2.0 * bytes / (mbest * 1e-3) / 1e9);
        }
    }
    CK(cudaGetLastError());
    return 0;
}
