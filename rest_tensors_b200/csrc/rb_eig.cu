// rb_eig.cu -- symmetric eigen-solvers of the reference's LAPACK wrappers:
//   _dsyev / lapack_dsyev        src/matrix/matrix_blas_lapack.rs:319-352, 775-797   (all eigenpairs, ascending)
//   lapack_dspevx                1075-1095                                            (packed upper input)
//   lapack_dspgvx / _dspgvx      1096-1147, 2123-2185                                 (A x = lambda B x, lowest num_orb pairs)
//   _power / lapack_power        599-652, 1004-1062                                   (A^p over the eigenvalues >= threshold)
//
// The reference calls LAPACK (tridiagonalisation + QR / bisection: sequential, host-shaped).  Here: a parallel one-sided
// Jacobi method (Hestenes), which is nothing but column rotations -- the access pattern the rest of this library is
// built around.  G starts as the symmetric matrix (shifted to be positive definite), n/2 disjoint column pairs are
// orthogonalised per launch (round-robin tournament ordering, n-1 launches per sweep), one CTA per pair: the two columns
// are read once (coalesced), their three inner products are reduced in a fixed order, and the rotated columns are
// written back from shared memory.  G and V (<= 2 n^2 doubles, 52 MB at n = 1800) stay in the 126 MB L2 across rounds.
// At convergence the columns of G are lambda_i v_i: the eigenvectors are the normalised columns (or the accumulated
// rotations for semi-definite input), the eigenvalues are Rayleigh quotients v_i^T S v_i of the ORIGINAL matrix (one DMMA
// GEMM + column dots), so the shift costs no accuracy.  Convergence: a sweep without a pair above
// |g_i.g_j| > sqrt(n) eps |g_i||g_j| (the dgesvj criterion).  Results agree with LAPACK to ~1e-13 relative to the norm;
// eigenvectors are defined up to sign (largest component made positive) and up to rotations inside degenerate spaces.
#include "rb_common.cuh"
#include <cooperative_groups.h>
#include <algorithm>
#include <cmath>
#include <memory>
#include <numeric>
#include <vector>

namespace {

constexpr int EIG_THREADS = 256;
constexpr int EIG_CACHE_MAX_N = 3060; // both columns of a pair (2 n doubles) + 192 B of static reduction scratch within the 48 KB non-opt-in limit
constexpr int EIG_MAX_SWEEPS = 60;

// g = symmetric expansion of one triangle of a (+ shift on the diagonal); v = identity (if v != NULL)
__global__ void __launch_bounds__(256) rb_eig_init_kernel(const double *__restrict__ a, i64 lda, int upper, double shift,
                                                          double *__restrict__ g, double *__restrict__ v, i64 n)
{
    const i64 total = n * n, stride = (i64)gridDim.x * blockDim.x;
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const i64 i = e % n, j = e / n;
        const bool in_tri = upper ? (i <= j) : (i >= j);
        double x = in_tri ? a[i + j * lda] : a[j + i * lda];
        if (i == j) x += shift;
        g[e] = x;
        if (v) v[e] = (i == j) ? 1.0 : 0.0;
    }
}

__global__ void __launch_bounds__(256) rb_eig_shift_diag_kernel(double *__restrict__ g, i64 n, double shift)
{
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) g[i + i * n] += shift;
}

// Jacobi rotation that makes two columns with squared norms a, b and inner product c orthogonal: tan(theta) is the smaller
// root of t^2 + 2 zeta t - 1 = 0, zeta = (b - a) / (2c).  Plain sqrt (hypot's overflow handling costs ~200 instructions):
// |zeta| > 1e150 only when c is negligible, and then t = 1 / (2 zeta) to working precision.
__device__ __forceinline__ void jacobi_rotation(double a, double b, double c, double &cs, double &sn)
{
    const double zeta = (b - a) / (2.0 * c);
    const double az = fabs(zeta);
    const double t = az > 1e150 ? 0.5 / zeta : copysign(1.0, zeta) / (az + sqrt(1.0 + zeta * zeta));
    cs = rsqrt(1.0 + t * t);
    cs = cs * (1.5 - 0.5 * (1.0 + t * t) * cs * cs); // one Newton step: rsqrt is ~1 ulp-ish, make it correctly rounded-ish
    sn = cs * t;
}

// fixed-order block reduction of three partial sums; every thread returns the totals
__device__ __forceinline__ void block_sum3(double &a, double &b, double &c, double *red)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
        c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { red[warp * 3 + 0] = a; red[warp * 3 + 1] = b; red[warp * 3 + 2] = c; }
    __syncthreads();
    a = 0.0; b = 0.0; c = 0.0;
#pragma unroll
    for (int w = 0; w < EIG_THREADS / 32; ++w) { a += red[w * 3 + 0]; b += red[w * 3 + 1]; c += red[w * 3 + 2]; }
    __syncthreads();
}

// per column: out[j] = sum_i |g[i,j]| (mode 0) or sqrt(sum_i g[i,j]^2) (mode 1); one CTA per column
__global__ void __launch_bounds__(EIG_THREADS) rb_eig_colstat_kernel(const double *__restrict__ g, i64 n, int mode,
                                                                    double *__restrict__ out)
{
    __shared__ double red[3 * EIG_THREADS / 32];
    for (i64 j = blockIdx.x; j < n; j += gridDim.x) {
        const double *col = g + j * n;
        double a = 0.0, b = 0.0, c = 0.0;
        for (i64 i = threadIdx.x; i < n; i += EIG_THREADS) {
            const double x = col[i];
            a += mode ? x * x : fabs(x);
        }
        block_sum3(a, b, c, red);
        if (threadIdx.x == 0) out[j] = mode ? sqrt(a) : a;
    }
}

// One round of the tournament: CTA k orthogonalises the column pair (i, j) of round `round`.
template <bool ACCUM_V, bool CACHE>
__global__ void __launch_bounds__(EIG_THREADS, 8) rb_jacobi_round_kernel(double *__restrict__ g, double *__restrict__ v, i64 n,
                                                                     i64 n_even, i64 round, double tol,
                                                                     unsigned long long *__restrict__ rotations)
{
    extern __shared__ double cols[]; // CACHE: [2][n]
    __shared__ double red[3 * EIG_THREADS / 32];
    const i64 m = n_even - 1, k = blockIdx.x;
    i64 i, j;
    if (k == 0) { i = round; j = m; }
    else { i = (round + k) % m; j = (round - k + m) % m; }
    if (i > j) { const i64 t = i; i = j; j = t; }
    if (j >= n) return; // odd n: this pair is the bye
    double *gi = g + i * n, *gj = g + j * n;
    double a = 0.0, b = 0.0, c = 0.0;
    for (i64 r = threadIdx.x; r < n; r += EIG_THREADS) {
        const double x = gi[r], y = gj[r];
        if (CACHE) { cols[r] = x; cols[n + r] = y; }
        a += x * x; b += y * y; c += x * y;
    }
    block_sum3(a, b, c, red);
    if (!(fabs(c) > tol * (sqrt(a) * sqrt(b)))) return; // already orthogonal (also: zero column, NaN)
    // A column that has collapsed to nothing (squared norm <= floor2 = (eps^2 |G|_F)^2, kept next to the rotation counter)
    // belongs to the null space of a rank-deficient matrix: what it holds is the rounding residue of earlier rotations, it
    // shrinks by a factor eps per sweep and never becomes orthogonal to the large columns in the RELATIVE sense of the
    // test above (emulated: a 40+23 bipartite matrix rotates 391 such pairs per sweep for ever; 11 sweeps with the floor).
    // The floor sits 16 orders of magnitude below working precision, so every eigenvalue that means anything keeps its
    // rotations.  floor2 = 0 in the shifted (definite) mode.
    const double floor2 = reinterpret_cast<const double *>(rotations)[1];
    if (a <= floor2 || b <= floor2) return;
    double cs, sn;
    jacobi_rotation(a, b, c, cs, sn);
    for (i64 r = threadIdx.x; r < n; r += EIG_THREADS) {
        const double x = CACHE ? cols[r] : gi[r], y = CACHE ? cols[n + r] : gj[r];
        gi[r] = cs * x - sn * y;
        gj[r] = sn * x + cs * y;
    }
    if (ACCUM_V) {
        double *vi = v + i * n, *vj = v + j * n;
        for (i64 r = threadIdx.x; r < n; r += EIG_THREADS) {
            const double x = vi[r], y = vj[r];
            vi[r] = cs * x - sn * y;
            vj[r] = sn * x + cs * y;
        }
    }
    if (threadIdx.x == 0) atomicAdd(rotations, 1ULL);
}

// ---- cluster-resident solver for small matrices ---------------------------------------------------------------------
// For n <= EIG_CLUSTER_MAX_N the whole working matrix lives in the DISTRIBUTED SHARED MEMORY of one thread-block cluster
// (column c in CTA c % C, slot c / C), and the complete solve -- every round of every sweep
// -- is ONE kernel: a warp takes a column pair into registers (<= 4 doubles per lane and column, local or remote shared
// memory alike), reduces the three inner products with shuffles, writes the rotated columns back, and cluster.sync()
// (~0.2 us) separates the rounds instead of a kernel boundary (a few microseconds per dependent graph node).  Convergence is decided uniformly by
// every CTA from the per-CTA rotation counts of the sweep (read through DSMEM).
namespace cg = cooperative_groups;
constexpr int EIG_CLUSTER_MAX_N = 128;    // 4 elements per lane and column.  Measured crossover against the graph-replayed round
                                          // kernels (tools/eig_small_probe.py): n = 64 1.26 vs 1.95 ms, 128 3.2 vs 3.6 ms, but
                                          // 200 7.7 vs 5.9 ms, 264 15.8 vs 8.2 ms -- remote shared memory traffic grows as n^2
                                          // per round on 8 SMs while the round kernels spread it over the chip's L2
constexpr int EIG_CLUSTER_THREADS = 512;  // 16 warps per CTA

__global__ void __launch_bounds__(EIG_CLUSTER_THREADS, 1)
rb_jacobi_cluster_kernel(double *__restrict__ g, int n, int n_even, double tol, int max_sweeps, int *__restrict__ result)
{
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ double cols[];            // this CTA's columns: [slots][n]
    __shared__ int rot_count[2];
    const int C = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = EIG_CLUSTER_THREADS / 32;
    const int slots = (n - rank + C - 1) / C;   // columns rank, rank + C, ...
    for (int sidx = 0; sidx < slots; ++sidx) {
        const double *src = g + (size_t)(rank + sidx * C) * n;
        for (int r = threadIdx.x; r < n; r += EIG_CLUSTER_THREADS) cols[(size_t)sidx * n + r] = src[r];
    }
    if (threadIdx.x < 2) rot_count[threadIdx.x] = 0;
    cluster.sync();
    const int m = n_even - 1, pairs = n_even / 2, total_warps = C * warps, gw = rank * warps + warp;
    int sweep = 0, converged = 0;
    for (; sweep < max_sweeps && !converged; ++sweep) {
        const int slot = sweep & 1;
        int my_rot = 0;
        for (int round = 0; round < m; ++round) {
            for (int k = gw; k < pairs; k += total_warps) {
                int i, j;
                if (k == 0) { i = round; j = m; }
                else { i = (round + k) % m; j = (round - k + m) % m; }
                if (i > j) { const int t = i; i = j; j = t; }
                if (j >= n) continue;           // odd n: the bye
                double *ci = cluster.map_shared_rank(cols + (size_t)(i / C) * n, i % C);
                double *cj = cluster.map_shared_rank(cols + (size_t)(j / C) * n, j % C);
                double x[EIG_CLUSTER_MAX_N / 32], y[EIG_CLUSTER_MAX_N / 32];
                double a = 0.0, b = 0.0, c = 0.0;
#pragma unroll
                for (int e = 0; e < EIG_CLUSTER_MAX_N / 32; ++e) {
                    const int r = lane + 32 * e;
                    x[e] = r < n ? ci[r] : 0.0;
                    y[e] = r < n ? cj[r] : 0.0;
                    a += x[e] * x[e]; b += y[e] * y[e]; c += x[e] * y[e];
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    a += __shfl_xor_sync(0xffffffffu, a, o);
                    b += __shfl_xor_sync(0xffffffffu, b, o);
                    c += __shfl_xor_sync(0xffffffffu, c, o);
                }
                if (!(fabs(c) > tol * (sqrt(a) * sqrt(b)))) continue;
                double cs, sn;
                jacobi_rotation(a, b, c, cs, sn);
#pragma unroll
                for (int e = 0; e < EIG_CLUSTER_MAX_N / 32; ++e) {
                    const int r = lane + 32 * e;
                    if (r < n) { ci[r] = cs * x[e] - sn * y[e]; cj[r] = sn * x[e] + cs * y[e]; }
                }
                ++my_rot;
            }
            cluster.sync();                     // every column of this round is written before the next pairing reads it
        }
        if (lane == 0 && my_rot) atomicAdd(&rot_count[slot], my_rot);
        cluster.sync();
        int total = 0;
        for (int q = 0; q < C; ++q) total += *cluster.map_shared_rank(&rot_count[slot], q);
        converged = total == 0;
        // the other slot is next written one sweep from now, after at least m more cluster barriers: safe to clear here
        if (threadIdx.x == 0) rot_count[slot ^ 1] = 0;
    }
    for (int sidx = 0; sidx < slots; ++sidx) {
        double *dst = g + (size_t)(rank + sidx * C) * n;
        for (int r = threadIdx.x; r < n; r += EIG_CLUSTER_THREADS) dst[r] = cols[(size_t)sidx * n + r];
    }
    if (rank == 0 && threadIdx.x == 0) { result[0] = sweep; result[1] = converged; }
    cluster.sync();                             // nobody leaves while a peer may still address its shared memory
}

// Returns RB_OK and sets *done = true when the cluster kernel ran (converged or not); *done = false: not applicable or the
// cluster could not be placed -> the caller uses the round-per-launch path.
int try_cluster_solve(rb_ctx *ctx, double *g, i64 n, double tol, int *result_dev, int *sweeps, bool *converged, bool *done)
{
    *done = false;
    if (n < 32 || n > EIG_CLUSTER_MAX_N) return RB_OK;
    if (const char *e = getenv("REST_B200_EIG_CLUSTER")) if (atoi(e) == 0) return RB_OK;
    const int csize = 8; // portable cluster size; 8 x 16 warps cover the <= 64 pairs of a round
    const size_t smem = (size_t)(rb_cdiv(n, csize) * n * 8);
    if (smem > 220 * 1024) return RB_OK;
    static bool attr_set[64] = {false};
    const int dev = ctx->device & 63;
    if (!attr_set[dev]) {
        if (cudaFuncSetAttribute(rb_jacobi_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) != cudaSuccess ||
            cudaFuncSetAttribute(rb_jacobi_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
            cudaGetLastError();
            return RB_OK;
        }
        attr_set[dev] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)csize, 1, 1);
    cfg.blockDim = dim3(EIG_CLUSTER_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int max_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&max_clusters, rb_jacobi_cluster_kernel, &cfg) != cudaSuccess || max_clusters < 1) {
        cudaGetLastError();
        return RB_OK; // this cluster shape cannot be placed on the device right now
    }
    const int ni = (int)n, ne = (int)(n + (n & 1)), ms = EIG_MAX_SWEEPS;
    if (cudaLaunchKernelEx(&cfg, rb_jacobi_cluster_kernel, g, ni, ne, tol, ms, result_dev) != cudaSuccess) {
        cudaGetLastError();
        return RB_OK;
    }
    ctx->launches++;
    int res[2] = {0, 0};
    RB_CUDA(cudaMemcpyAsync(res, result_dev, 8, cudaMemcpyDeviceToHost, ctx->stream));
    RB_CUDA(cudaStreamSynchronize(ctx->stream));
    *sweeps = res[0]; *converged = res[1] != 0; *done = true;
    return RB_OK;
}

// v[:, j] = g[:, j] / norm[j]
__global__ void __launch_bounds__(256) rb_eig_normalize_kernel(const double *__restrict__ g, const double *__restrict__ norm,
                                                               double *__restrict__ v, i64 n)
{
    const i64 total = n * n, stride = (i64)gridDim.x * blockDim.x;
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const double d = norm[e / n];
        v[e] = d > 0.0 ? g[e] / d : 0.0;
    }
}

// z[:, c] = sign * v[:, perm[c]] for c < ncols, sign chosen so that the component of largest magnitude (lowest index on
// ties) is positive; one CTA per output column
__global__ void __launch_bounds__(EIG_THREADS) rb_eig_permute_kernel(const double *__restrict__ v, i64 n,
                                                                    const i64 *__restrict__ perm, double *__restrict__ z,
                                                                    i64 ldz, i64 ncols)
{
    __shared__ double bestv[EIG_THREADS];
    __shared__ i64 besti[EIG_THREADS];
    for (i64 c = blockIdx.x; c < ncols; c += gridDim.x) {
        const double *src = v + perm[c] * n;
        double bv = -1.0;
        i64 bi = 0;
        for (i64 r = threadIdx.x; r < n; r += EIG_THREADS) {
            const double x = fabs(src[r]);
            if (x > bv) { bv = x; bi = r; }
        }
        bestv[threadIdx.x] = bv; besti[threadIdx.x] = bi;
        __syncthreads();
        for (int o = EIG_THREADS / 2; o > 0; o >>= 1) {
            if (threadIdx.x < o) {
                const double ov = bestv[threadIdx.x + o];
                const i64 oi = besti[threadIdx.x + o];
                if (ov > bestv[threadIdx.x] || (ov == bestv[threadIdx.x] && oi < besti[threadIdx.x])) {
                    bestv[threadIdx.x] = ov; besti[threadIdx.x] = oi;
                }
            }
            __syncthreads();
        }
        const double sgn = (n > 0 && src[besti[0]] < 0.0) ? -1.0 : 1.0;
        __syncthreads();
        for (i64 r = threadIdx.x; r < n; r += EIG_THREADS) z[r + c * ldz] = sgn * src[r];
    }
}

int grid_for(rb_ctx *ctx, i64 total, int per_block)
{
    i64 blocks = rb_cdiv(total, per_block);
    const i64 cap = (i64)ctx->num_sms * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

// A round is a single dependent step, so its duration is (waves of CTAs) x (one CTA's latency chain): the shared-memory
// column cache is used only when it does not cost a wave.  Resident CTAs per SM: 8 by threads and registers (launch bounds:
// 32 registers); with the cache also by shared memory (2 n doubles + 1 KB reserved each, maximum carve-out requested).
// n = 1800 (900 pairs): 7 x 148 = 1036 slots hold the round in one wave, where the 6 CTAs per SM of the first build (40
// registers; ncu: 1.01 waves) ran a second, nearly empty one.
bool eig_use_cache(rb_ctx *ctx, i64 n, i64 n_even)
{
    if (n > EIG_CACHE_MAX_N) return false;
    static bool carve_set[64] = {false};
    const int dev = ctx->device & 63;
    if (!carve_set[dev]) {
        cudaFuncSetAttribute(rb_jacobi_round_kernel<false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(rb_jacobi_round_kernel<true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaGetLastError();
        carve_set[dev] = true;
    }
    const i64 pairs = n_even / 2;
    i64 occ_cache = (i64)(227 * 1024) / (2 * n * 8 + 1024);
    if (occ_cache > 8) occ_cache = 8;
    if (occ_cache < 1) occ_cache = 1;
    return rb_cdiv(pairs, occ_cache * ctx->num_sms) <= rb_cdiv(pairs, (i64)8 * ctx->num_sms);
}

template <bool ACCUM_V>
int launch_round(rb_ctx *ctx, double *g, double *v, i64 n, i64 n_even, i64 round, double tol, unsigned long long *rot)
{
    const unsigned blocks = (unsigned)(n_even / 2);
    if (eig_use_cache(ctx, n, n_even))
        rb_jacobi_round_kernel<ACCUM_V, true><<<blocks, EIG_THREADS, (size_t)(2 * n * 8), ctx->stream>>>(g, v, n, n_even, round, tol, rot);
    else
        rb_jacobi_round_kernel<ACCUM_V, false><<<blocks, EIG_THREADS, 0, ctx->stream>>>(g, v, n, n_even, round, tol, rot);
    RB_LAUNCHED(ctx);
    return RB_OK;
}

// A whole sweep as an executable graph: memset of the rotation counter, then the n-1 rounds as a chain of kernel nodes.
struct SweepGraph {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    ~SweepGraph()
    {
        if (exec) cudaGraphExecDestroy(exec);
        if (graph) cudaGraphDestroy(graph);
    }
};

bool build_sweep_graph(rb_ctx *ctx, SweepGraph &sg, bool psd, double *g, double *v, i64 n, i64 n_even, double tol,
                       unsigned long long *rot)
{
    if (const char *e = getenv("REST_B200_EIG_GRAPH")) if (atoi(e) == 0) return false;
    if (cudaGraphCreate(&sg.graph, 0) != cudaSuccess) { cudaGetLastError(); return false; }
    cudaMemsetParams mp = {};
    mp.dst = rot; mp.value = 0; mp.elementSize = 4; mp.width = 2; mp.height = 1; mp.pitch = 8;
    cudaGraphNode_t prev = nullptr;
    if (cudaGraphAddMemsetNode(&prev, sg.graph, nullptr, 0, &mp) != cudaSuccess) { cudaGetLastError(); return false; }
    const bool cache = eig_use_cache(ctx, n, n_even);
    void *fn = psd ? (cache ? (void *)rb_jacobi_round_kernel<true, true> : (void *)rb_jacobi_round_kernel<true, false>)
                   : (cache ? (void *)rb_jacobi_round_kernel<false, true> : (void *)rb_jacobi_round_kernel<false, false>);
    for (i64 round = 0; round < n_even - 1; ++round) {
        i64 round_arg = round;
        void *args[7] = {&g, &v, &n, &n_even, &round_arg, &tol, &rot};
        cudaKernelNodeParams kp = {};
        kp.func = fn;
        kp.gridDim = dim3((unsigned)(n_even / 2), 1, 1);
        kp.blockDim = dim3(EIG_THREADS, 1, 1);
        kp.sharedMemBytes = cache ? (unsigned)(2 * n * 8) : 0u;
        kp.kernelParams = args; // copied by the call
        kp.extra = nullptr;
        cudaGraphNode_t node = nullptr;
        if (cudaGraphAddKernelNode(&node, sg.graph, &prev, 1, &kp) != cudaSuccess) { cudaGetLastError(); return false; }
        prev = node;
    }
    if (cudaGraphInstantiate(&sg.exec, sg.graph, 0) != cudaSuccess) { cudaGetLastError(); sg.exec = nullptr; return false; }
    (void)ctx;
    return true;
}

// Building and instantiating the n-1 node chain costs milliseconds of host time (more when the host cores are busy), and an
// SCF loop diagonalises matrices of one size over and over in the same workspace: the instantiated graphs are kept per
// context, keyed by everything the kernel nodes captured (shape, mode, buffers, tolerance); at most 4, least recently
// used replaced.
struct EigGraphEntry {
    i64 n = 0;
    bool psd = false;
    double *g = nullptr, *v = nullptr;
    double tol = 0.0;
    unsigned long long *rot = nullptr;
    unsigned stamp = 0;
    SweepGraph sg;
};
struct EigCache {
    std::vector<std::unique_ptr<EigGraphEntry>> entries;
    unsigned clock = 0;
};

SweepGraph *get_sweep_graph(rb_ctx *ctx, bool psd, double *g, double *v, i64 n, i64 n_even, double tol, unsigned long long *rot)
{
    if (!ctx->eig_cache) ctx->eig_cache = new EigCache();
    EigCache *cache = (EigCache *)ctx->eig_cache;
    ++cache->clock;
    for (auto &e : cache->entries)
        if (e->n == n && e->psd == psd && e->g == g && e->v == v && e->tol == tol && e->rot == rot) {
            e->stamp = cache->clock;
            return &e->sg;
        }
    std::unique_ptr<EigGraphEntry> ent(new EigGraphEntry());
    if (!build_sweep_graph(ctx, ent->sg, psd, g, v, n, n_even, tol, rot)) return nullptr;
    ent->n = n; ent->psd = psd; ent->g = g; ent->v = v; ent->tol = tol; ent->rot = rot; ent->stamp = cache->clock;
    if (cache->entries.size() >= 4) {
        size_t old = 0;
        for (size_t i = 1; i < cache->entries.size(); ++i) if (cache->entries[i]->stamp < cache->entries[old]->stamp) old = i;
        cudaStreamSynchronize(ctx->stream); // the evicted graph may still be running
        cache->entries.erase(cache->entries.begin() + (long)old);
    }
    cache->entries.push_back(std::move(ent));
    return &cache->entries.back()->sg;
}

// Workspace of one solve, in doubles: G, V, W (n^2 each) + lam, stat (n each) + perm (n int64) + the rotation counter.
i64 eig_work_elems(i64 n) { return 3 * n * n + 3 * n + 8; }

// Eigen-decomposition of the full symmetric n x n matrix s (dense, ld = n; not modified).  psd: no shift and accumulated
// rotations (accurate small eigenvalues of positive semi-definite input); else Gershgorin shift and normalised columns.
// lam_host[n] ascending; z (device, ldz) receives the first ncols eigenvectors (NULL: none).  Synchronises the stream.
int jacobi_eig(rb_ctx *ctx, i64 n, const double *s, bool psd, double *work, std::vector<double> &lam_host, double *z, i64 ldz,
               i64 ncols, int *sweeps_out, bool *psd_ok = nullptr)
{
    lam_host.assign((size_t)n, 0.0);
    if (sweeps_out) *sweeps_out = 0;
    if (psd_ok) *psd_ok = true;
    if (n == 0) return RB_OK;
    double *g = work, *v = g + n * n, *w = v + n * n, *lam = w + n * n, *stat = lam + n;
    i64 *perm = (i64 *)(stat + n);
    unsigned long long *rot = (unsigned long long *)(perm + n);
    std::vector<double> host((size_t)n);
    double shift = 0.0;
    rb_eig_init_kernel<<<grid_for(ctx, n * n, 256), 256, 0, ctx->stream>>>(s, n, 1, 0.0, g, psd ? v : nullptr, n);
    RB_LAUNCHED(ctx);
    if (!psd) { // Gershgorin: every eigenvalue lies in [-R, R], R = max column sum of |s|; G = S + 1.5 R I has cond <= 5
        rb_eig_colstat_kernel<<<grid_for(ctx, n, 1), EIG_THREADS, 0, ctx->stream>>>(g, n, 0, stat);
        RB_LAUNCHED(ctx);
        RB_CUDA(cudaMemcpyAsync(host.data(), stat, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
        RB_CUDA(cudaStreamSynchronize(ctx->stream));
        double r = 0.0;
        for (double x : host) {
            if (!(x == x) || std::isinf(x)) { rb_set_error("eigen-solver: the matrix holds non-finite values"); return RB_ERR_INVALID; }
            if (x > r) r = x;
        }
        shift = r > 0.0 ? 1.5 * r : 1.0; // zero matrix: G = I, every vector is an eigenvector
        if (shift > 0.0) {
            rb_eig_shift_diag_kernel<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(g, n, shift);
            RB_LAUNCHED(ctx);
        }
    }
    const i64 n_even = n + (n & 1);
    const double tol = std::sqrt((double)n) * 1.1102230246251565e-16;
    // collapsed-column floor of the unshifted mode (see rb_jacobi_round_kernel): (eps^2 |G|_F)^2, |G|_F is invariant under
    // the rotations.  It lives in device memory (rot[1]) so that the cached sweep graphs stay valid from matrix to matrix.
    double floor2 = 0.0;
    if (psd) {
        rb_eig_colstat_kernel<<<grid_for(ctx, n, 1), EIG_THREADS, 0, ctx->stream>>>(g, n, 1, stat);
        RB_LAUNCHED(ctx);
        RB_CUDA(cudaMemcpyAsync(host.data(), stat, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
        RB_CUDA(cudaStreamSynchronize(ctx->stream));
        double fro2 = 0.0;
        for (double x : host) {
            if (!(x == x) || std::isinf(x)) { rb_set_error("eigen-solver: the matrix holds non-finite values"); return RB_ERR_INVALID; }
            fro2 += x * x;
        }
        const double e2 = 2.220446049250313e-16 * 2.220446049250313e-16;
        floor2 = e2 * e2 * fro2;
    }
    RB_CUDA(cudaMemcpyAsync(rot + 1, &floor2, 8, cudaMemcpyHostToDevice, ctx->stream));
    RB_CUDA(cudaStreamSynchronize(ctx->stream)); // floor2 is a stack variable
    int sweeps = 0;
    bool converged = n < 2;
    if (!psd && !converged) { // small matrices: the whole solve as one cluster-resident kernel
        bool done = false;
        RB_TRY(try_cluster_solve(ctx, g, n, tol, (int *)rot, &sweeps, &converged, &done));
        if (done && !converged) sweeps = EIG_MAX_SWEEPS; // fall through to the error below
    }
    // One sweep = n-1 dependent launches of a few microseconds each: issued one by one the host's launch cost (~5 us)
    // bounds the sweep, so the sweep is built once as a CUDA graph (a chain of kernel nodes, one per round) and replayed.
    SweepGraph *graph = (!converged && n_even - 1 >= 16) ? get_sweep_graph(ctx, psd, g, v, n, n_even, tol, rot) : nullptr;
    bool use_graph = graph != nullptr;
    for (; sweeps < EIG_MAX_SWEEPS && !converged; ++sweeps) {
        if (use_graph && cudaGraphLaunch(graph->exec, ctx->stream) != cudaSuccess) { // e.g. a stream that takes no graphs
            cudaGetLastError();
            use_graph = false;
        }
        if (use_graph) {
            ctx->launches += n_even - 1;
        } else {
            RB_CUDA(cudaMemsetAsync(rot, 0, 8, ctx->stream));
            for (i64 round = 0; round < n_even - 1; ++round) {
                if (psd) RB_TRY(launch_round<true>(ctx, g, v, n, n_even, round, tol, rot));
                else RB_TRY(launch_round<false>(ctx, g, v, n, n_even, round, tol, rot));
            }
        }
        unsigned long long nrot = 0;
        RB_CUDA(cudaMemcpyAsync(&nrot, rot, 8, cudaMemcpyDeviceToHost, ctx->stream));
        RB_CUDA(cudaStreamSynchronize(ctx->stream));
        converged = nrot == 0;
    }
    if (sweeps_out) *sweeps_out = sweeps;
    if (!converged) {
        rb_set_error("eigen-solver: Jacobi sweeps did not converge in %d sweeps (n = %lld)", EIG_MAX_SWEEPS, (long long)n);
        return RB_ERR_CUDA;
    }
    if (!psd) { // eigenvectors = normalised columns of G (= lambda_i v_i with lambda_i >= 0.5 R > 0)
        rb_eig_colstat_kernel<<<grid_for(ctx, n, 1), EIG_THREADS, 0, ctx->stream>>>(g, n, 1, stat);
        RB_LAUNCHED(ctx);
        rb_eig_normalize_kernel<<<grid_for(ctx, n * n, 256), 256, 0, ctx->stream>>>(g, stat, v, n);
        RB_LAUNCHED(ctx);
    }
    // Rayleigh quotients with the original matrix: W = S V, lam_i = v_i . w_i
    RB_TRY(rb_gemm_core(ctx, false, false, n, n, n, 1.0, s, n, 0, v, n, 0, 0.0, w, n, 0, 1, 0));
    RB_TRY(rb_einsum_ip_ip(ctx, v, n, w, n, lam, n, n));
    RB_CUDA(cudaMemcpyAsync(host.data(), lam, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (psd) { // semi-definite input: |g_i| = lambda_i carries the relative accuracy of small eigenvalues, the quotient does not
        rb_eig_colstat_kernel<<<grid_for(ctx, n, 1), EIG_THREADS, 0, ctx->stream>>>(g, n, 1, stat);
        RB_LAUNCHED(ctx);
    }
    RB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (psd) {
        // Unshifted one-sided Jacobi diagonalises S^2: it only sees |lambda|, and the eigenvectors of +s / -s pairs may stay
        // mixed (e.g. [[0,1],[1,0]]: the columns are orthogonal from the start).  The converged columns are accepted as
        // eigenpairs of a semi-definite S only if EVERY Rayleigh quotient v_i^T S v_i agrees with its column norm |g_i|
        // to rounding (both are lambda_i >= 0 then; a mixed or negative pair gives a quotient below its norm).  Otherwise
        // the caller repeats the solve with the shift, which is valid for any symmetric matrix.
        std::vector<double> norms((size_t)n);
        RB_CUDA(cudaMemcpyAsync(norms.data(), stat, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
        RB_CUDA(cudaStreamSynchronize(ctx->stream));
        double hi = 0.0, mismatch = 0.0;
        for (i64 i = 0; i < n; ++i) {
            hi = std::max(hi, norms[(size_t)i]);
            mismatch = std::max(mismatch, std::fabs(host[(size_t)i] - norms[(size_t)i]));
        }
        const bool ok = mismatch <= 64.0 * (double)n * 2.220446049250313e-16 * hi;
        if (psd_ok) *psd_ok = ok;
        if (!ok) return RB_OK; // lam_host / z are not valid; see psd_ok
        host = norms;          // |g_i| carries the relative accuracy of small eigenvalues
    }
    std::vector<i64> order((size_t)n);
    std::iota(order.begin(), order.end(), (i64)0);
    std::stable_sort(order.begin(), order.end(), [&](i64 x, i64 y) { return host[(size_t)x] < host[(size_t)y]; });
    for (i64 c = 0; c < n; ++c) lam_host[(size_t)c] = host[(size_t)order[(size_t)c]];
    if (z && ncols > 0) {
        RB_CUDA(cudaMemcpyAsync(perm, order.data(), (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
        rb_eig_permute_kernel<<<grid_for(ctx, ncols, 1), EIG_THREADS, 0, ctx->stream>>>(v, n, perm, z, ldz, ncols);
        RB_LAUNCHED(ctx);
        RB_CUDA(cudaStreamSynchronize(ctx->stream)); // `order` must outlive the upload
    }
    return RB_OK;
}

// positive semi-definite path with a fallback: if the Rayleigh quotients do not confirm the column norms as eigenvalues
// (indefinite input), the solve is repeated with the shift
int jacobi_eig_psd(rb_ctx *ctx, i64 n, const double *s, double *work, std::vector<double> &lam, double *z, i64 ldz, i64 ncols)
{
    bool ok = true;
    RB_TRY(jacobi_eig(ctx, n, s, true, work, lam, z, ldz, ncols, nullptr, &ok));
    if (!ok) RB_TRY(jacobi_eig(ctx, n, s, false, work, lam, z, ldz, ncols, nullptr));
    return RB_OK;
}

} // namespace

extern "C" int rb_dsyev(rb_ctx *ctx, char jobz, char uplo, int n_, const double *a, int64_t lda, double *w, double *z,
                        int64_t ldz)
{
    RB_REQUIRE(ctx, "rb_dsyev: ctx is NULL");
    RB_NO_CAPTURE(ctx, "rb_dsyev");
    RB_REQUIRE(jobz == 'V' || jobz == 'v' || jobz == 'N' || jobz == 'n', "rb_dsyev: jobz must be 'V' or 'N'");
    RB_REQUIRE(rb_is_u(uplo) || rb_is_l(uplo), "rb_dsyev: uplo must be 'U' or 'L'");
    RB_REQUIRE(n_ >= 0, "rb_dsyev: negative dimension");
    const i64 n = n_;
    if (n == 0) return RB_OK;
    const bool want_z = jobz == 'V' || jobz == 'v';
    RB_REQUIRE(a && w && lda >= n && (!want_z || (z && ldz >= n)), "rb_dsyev: NULL buffer or leading dimension too small");
    RB_CUDA(cudaSetDevice(ctx->device));
    void *ws;
    RB_TRY(rb_ws_reserve(ctx, 0, (n * n + eig_work_elems(n)) * 8, &ws));
    double *s = (double *)ws, *work = s + n * n;
    rb_eig_init_kernel<<<grid_for(ctx, n * n, 256), 256, 0, ctx->stream>>>(a, lda, rb_is_u(uplo) ? 1 : 0, 0.0, s, nullptr, n);
    RB_LAUNCHED(ctx);
    std::vector<double> lam;
    RB_TRY(jacobi_eig(ctx, n, s, false, work, lam, want_z ? z : nullptr, ldz, want_z ? n : 0, nullptr));
    RB_CUDA(cudaMemcpyAsync(w, lam.data(), (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    RB_CUDA(cudaStreamSynchronize(ctx->stream));
    return RB_OK;
}

extern "C" int rb_dspev(rb_ctx *ctx, int n_, const double *ap, double *w, double *z, int64_t ldz)
{
    RB_REQUIRE(ctx, "rb_dspev: ctx is NULL");
    RB_NO_CAPTURE(ctx, "rb_dspev");
    RB_REQUIRE(n_ >= 0, "rb_dspev: negative dimension");
    const i64 n = n_;
    if (n == 0) return RB_OK;
    RB_REQUIRE(ap && w && (!z || ldz >= n), "rb_dspev: NULL buffer or ldz too small");
    RB_CUDA(cudaSetDevice(ctx->device));
    void *ws;
    RB_TRY(rb_ws_reserve(ctx, 0, (n * n + eig_work_elems(n)) * 8, &ws));
    double *s = (double *)ws, *work = s + n * n;
    RB_TRY(rb_unpack_upper(ctx, ap, n, s)); // full symmetric matrix, both triangles
    std::vector<double> lam;
    RB_TRY(jacobi_eig(ctx, n, s, false, work, lam, z, ldz, z ? n : 0, nullptr));
    RB_CUDA(cudaMemcpyAsync(w, lam.data(), (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    RB_CUDA(cudaStreamSynchronize(ctx->stream));
    return RB_OK;
}

extern "C" int rb_dspgv(rb_ctx *ctx, int n_, const double *ap, const double *bp, int m_, double *w, double *z, int64_t ldz)
{
    RB_REQUIRE(ctx, "rb_dspgv: ctx is NULL");
    RB_NO_CAPTURE(ctx, "rb_dspgv");
    RB_REQUIRE(n_ >= 0 && m_ >= 0 && m_ <= n_, "rb_dspgv: bad dimensions (n = %d, num_orb = %d)", n_, m_);
    const i64 n = n_, m = m_;
    if (n == 0 || m == 0) return RB_OK;
    RB_REQUIRE(ap && bp && w && z && ldz >= n, "rb_dspgv: NULL buffer or ldz too small");
    RB_CUDA(cudaSetDevice(ctx->device));
    void *ws;
    RB_TRY(rb_ws_reserve(ctx, 0, (5 * n * n + n + eig_work_elems(n)) * 8, &ws));
    double *a = (double *)ws, *b = a + n * n, *us = b + n * n, *t1 = us + n * n, *y = t1 + n * n, *scale = y + n * n,
           *work = scale + n;
    RB_TRY(rb_unpack_upper(ctx, ap, n, a));
    RB_TRY(rb_unpack_upper(ctx, bp, n, b));
    // B = U D U^T (positive definite); Us = U D^-1/2 makes Us^T B Us = I
    std::vector<double> d;
    RB_TRY(jacobi_eig_psd(ctx, n, b, work, d, us, n, n));
    // positive definite to working precision (what the reference's LAPACK Cholesky inside dspgvx demands; it panics otherwise)
    RB_REQUIRE(d[0] > 0.0 && d[0] > (double)n * 2.220446049250313e-16 * d[(size_t)n - 1],
               "rb_dspgv: the overlap matrix is not positive definite (eigenvalues %.3e .. %.3e)", d[0], d[(size_t)n - 1]);
    std::vector<double> sc((size_t)n);
    for (i64 i = 0; i < n; ++i) sc[(size_t)i] = 1.0 / std::sqrt(d[(size_t)i]);
    RB_CUDA(cudaMemcpyAsync(scale, sc.data(), (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    RB_TRY(rb_einsum_ij_j(ctx, us, n, scale, t1, n, n, n)); // t1 = Us (scaled copy)
    RB_CUDA(cudaStreamSynchronize(ctx->stream));
    // A' = Us^T A Us
    RB_TRY(rb_gemm_core(ctx, false, false, n, n, n, 1.0, a, n, 0, t1, n, 0, 0.0, y, n, 0, 1, 0));   // y = A Us
    RB_TRY(rb_gemm_core(ctx, true, false, n, n, n, 1.0, t1, n, 0, y, n, 0, 0.0, b, n, 0, 1, 1));    // b = upper(Us^T A Us)
    RB_TRY(rb_symmetrize(ctx, b, n, n, true));
    std::vector<double> lam;
    RB_TRY(jacobi_eig(ctx, n, b, false, work, lam, y, n, m, nullptr));                                // y[:, :m] = eigenvectors of A'
    RB_TRY(rb_gemm_core(ctx, false, false, n, m, n, 1.0, t1, n, 0, y, n, 0, 0.0, z, ldz, 0, 1, 0));  // z = Us Y
    RB_CUDA(cudaMemcpyAsync(w, lam.data(), (size_t)m * 8, cudaMemcpyHostToDevice, ctx->stream));
    RB_CUDA(cudaStreamSynchronize(ctx->stream));
    return RB_OK;
}

extern "C" int rb_matrix_power(rb_ctx *ctx, int n_, const double *a, int64_t lda, double p, double threshold, double *out,
                               int64_t ldo, int *n_nonsingular)
{
    RB_REQUIRE(ctx, "rb_matrix_power: ctx is NULL");
    RB_NO_CAPTURE(ctx, "rb_matrix_power");
    RB_REQUIRE(n_ >= 0, "rb_matrix_power: negative dimension");
    const i64 n = n_;
    if (n_nonsingular) *n_nonsingular = 0;
    if (n == 0) return RB_OK;
    RB_REQUIRE(a && out && lda >= n && ldo >= n, "rb_matrix_power: NULL buffer or leading dimension too small");
    RB_CUDA(cudaSetDevice(ctx->device));
    void *ws;
    RB_TRY(rb_ws_reserve(ctx, 0, (3 * n * n + n + eig_work_elems(n)) * 8, &ws));
    double *s = (double *)ws, *vz = s + n * n, *vs = vz + n * n, *scale = vs + n * n, *work = scale + n;
    // dsyev(.., 'L', ..) in the reference: the lower triangle is the one that is read
    rb_eig_init_kernel<<<grid_for(ctx, n * n, 256), 256, 0, ctx->stream>>>(a, lda, 0, 0.0, s, nullptr, n);
    RB_LAUNCHED(ctx);
    std::vector<double> lam;
    RB_TRY(jacobi_eig_psd(ctx, n, s, work, lam, vz, n, n));
    std::vector<double> sc((size_t)n);
    int kept = 0;
    for (i64 i = 0; i < n; ++i) {
        const double ev = lam[(size_t)i];
        if (ev >= threshold) { sc[(size_t)i] = std::pow(std::sqrt(ev), p); ++kept; }
        else sc[(size_t)i] = 0.0;
    }
    if (n_nonsingular) *n_nonsingular = kept;
    RB_CUDA(cudaMemcpyAsync(scale, sc.data(), (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    RB_TRY(rb_einsum_ij_j(ctx, vz, n, scale, vs, n, n, n));
    RB_CUDA(cudaStreamSynchronize(ctx->stream));
    RB_TRY(rb_gemm_core(ctx, false, true, n, n, n, 1.0, vs, n, 0, vs, n, 0, 0.0, out, ldo, 0, 1, 1));
    return rb_symmetrize(ctx, out, n, ldo, true);
}

void rb_eig_cache_free(rb_ctx *ctx)
{
    if (!ctx || !ctx->eig_cache) return;
    delete (EigCache *)ctx->eig_cache;
    ctx->eig_cache = nullptr;
}
