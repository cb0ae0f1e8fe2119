// rb_eri.cu -- ERIFold4 (reference src/eri.rs:170-373): the four-index integrals (ij|kl) with both pairs
// folded, stored as a column-major [npair_ij, npair_kl] matrix, packed index j(j+1)/2 + i with i <= j (src/index.rs:227-233).
// The hot member is the chunk scatter that moves a dense shell-quartet block [len_i, len_j, len_k, len_l] (as libcint
// returns it) into the folded tensor:
//   chunk_copy_from_local_erifull   (eri.rs:266-305): element (i, j, k, l) is kept iff k <= l and i <= j
//   chunk_copy_from_a_full_vector   (eri.rs:308-372): range_i.start <  range_j.start -> every (i, j) of the block (k <= l),
//                                                     range_i.start == range_j.start -> local ii <= jj (k <= l),
//                                                     otherwise nothing is copied
// Bit-exact data movement, HBM-bound: one thread per source element, reads coalesced along i, writes coalesced inside each
// packed run  dst[(l(l+1)/2 + k) * ld + j(j+1)/2 + i].
#include "rb_common.cuh"

__global__ void __launch_bounds__(256) rb_erifold4_scatter_kernel(double *__restrict__ dst, i64 ld, const double *__restrict__ buf, i64 i0,
                                                                  i64 li, i64 j0, i64 lj, i64 k0, i64 lk, i64 l0, i64 ll, int mode)
{
    const i64 total = li * lj * lk * ll, stride = (i64)gridDim.x * blockDim.x;
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const i64 ii = e % li, r1 = e / li;
        const i64 jj = r1 % lj, r2 = r1 / lj;
        const i64 kk = r2 % lk, lx = r2 / lk;
        const i64 i = i0 + ii, j = j0 + jj, k = k0 + kk, l = l0 + lx;
        if (k > l) continue;
        bool keep;
        if (mode == 0) keep = i <= j;
        else keep = (i0 < j0) || (i0 == j0 && ii <= jj);
        if (!keep) continue;
        dst[(l * (l + 1) / 2 + k) * ld + j * (j + 1) / 2 + i] = buf[e];
    }
}

// launch without bounds checks: `dst` may be a virtual origin (the host form scatters into a window of the tensor)
int rb_erifold4_scatter(rb_ctx *ctx, double *dst, i64 ld, const double *buf, i64 i0, i64 li, i64 j0, i64 lj, i64 k0, i64 lk, i64 l0,
                        i64 ll, int mode)
{
    const i64 total = li * lj * lk * ll;
    if (total <= 0) return RB_OK;
    i64 blocks = rb_cdiv(total, 256);
    const i64 cap = (i64)ctx->num_sms * 16;
    if (blocks > cap) blocks = cap;
    rb_erifold4_scatter_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(dst, ld, buf, i0, li, j0, lj, k0, lk, l0, ll, mode);
    RB_LAUNCHED(ctx);
    return RB_OK;
}

// Largest destination row / column the block touches (for the bounds check the reference does with `unwrap()` on every slice)
static bool erifold4_block_fits(i64 size0, i64 size1, i64 i0, i64 li, i64 j0, i64 lj, i64 k0, i64 lk, i64 l0, i64 ll, int mode)
{
    if (li <= 0 || lj <= 0 || lk <= 0 || ll <= 0) return true;
    if (mode == 1 && i0 > j0) return true; // nothing is copied
    const i64 jmax = j0 + lj - 1, lmax = l0 + ll - 1;
    i64 imax = i0 + li - 1;
    if (mode == 0 || i0 == j0) { if (imax > jmax) imax = jmax; }
    i64 kmax = k0 + lk - 1;
    if (kmax > lmax) kmax = lmax;
    if (k0 > lmax) return true; // every k > l
    const i64 row = jmax * (jmax + 1) / 2 + imax, col = lmax * (lmax + 1) / 2 + kmax;
    return row < size0 && col < size1;
}

extern "C" int rb_erifold4_chunk_copy(rb_ctx *ctx, double *eri, int64_t size0, int64_t size1, int64_t ld, int i0, int li, int j0,
                                      int lj, int k0, int lk, int l0, int ll, const double *buf, int mode)
{
    RB_REQUIRE(ctx, "rb_erifold4_chunk_copy: ctx is NULL");
    RB_REQUIRE(mode == 0 || mode == 1, "rb_erifold4_chunk_copy: mode must be 0 (local erifull) or 1 (full vector)");
    RB_REQUIRE(size0 >= 0 && size1 >= 0 && ld >= size0, "rb_erifold4_chunk_copy: bad tensor shape");
    RB_REQUIRE(i0 >= 0 && j0 >= 0 && k0 >= 0 && l0 >= 0 && li >= 0 && lj >= 0 && lk >= 0 && ll >= 0,
               "rb_erifold4_chunk_copy: negative range");
    const i64 total = (i64)li * lj * lk * ll;
    if (total == 0) return RB_OK;
    RB_REQUIRE(eri && buf, "rb_erifold4_chunk_copy: NULL buffer");
    RB_REQUIRE(erifold4_block_fits(size0, size1, i0, li, j0, lj, k0, lk, l0, ll, mode),
               "rb_erifold4_chunk_copy: the block reaches outside the folded tensor [%lld, %lld]", (long long)size0, (long long)size1);
    RB_CUDA(cudaSetDevice(ctx->device));
    return rb_erifold4_scatter(ctx, eri, ld, buf, i0, li, j0, lj, k0, lk, l0, ll, mode);
}
