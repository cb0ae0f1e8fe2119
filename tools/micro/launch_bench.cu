// launch_bench.cu -- does the kernel-parameter size matter for the launch rate?  Empty kernels with 64 B / 512 B / 2.5 KB / 4 KB of
// __grid_constant__ parameters, 148 x 384 threads and 198 KB of dynamic shared memory (the GEMM kernel's launch shape), back to back.
#include <cuda_runtime.h>
#include <cstdio>
template <int N> struct P { unsigned char b[N]; };
template <int N> __global__ void __launch_bounds__(384, 1) k(const __grid_constant__ P<N> p, int *out)
{
    extern __shared__ unsigned char sm[];
    if (threadIdx.x == 0 && p.b[N - 1] == 77) out[0] = sm[0];
}
template <int N> void run(int smem, const char *tag)
{
    P<N> p = {};
    int *out; cudaMalloc(&out, 4);
    cudaFuncSetAttribute(k<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 100; ++i) k<N><<<148, 384, smem>>>(p, out);
    cudaDeviceSynchronize();
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        for (int i = 0; i < 1000; ++i) k<N><<<148, 384, smem>>>(p, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    printf("%-28s params %4d B, smem %6d B: %.2f us per launch\n", tag, N, smem, best);
    cudaFree(out);
}
int main()
{
    run<64>(0, "empty kernel");
    run<64>(198 * 1024, "empty kernel");
    run<512>(198 * 1024, "GEMM params (round 1)");
    run<2560>(198 * 1024, "GEMM params + SkTables");
    run<4000>(198 * 1024, "near the 4 KB limit");
    return 0;
}
