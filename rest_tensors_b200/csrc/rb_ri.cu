// rb_ri.cu -- the RI three-centre contractions over one rank's P-shard ri3ao[nb, nb, nx] (device-resident).
//
//   ao2mo  (reference src/ri.rs:356-408 -> restmatr.f90:158-194)
//       ri3mo[P,a,b] = sum_mu C_L[mu,a] sum_nu A[mu,nu,P] C_R[nu,b]
//     Two DMMA GEMMs per P-chunk, both with K-major (TMA no-swizzle) operands and column-major output:
//       (1) W[(nu,P), a] = sum_mu A[mu,(nu,P)] C_L[mu,a]         -- ONE 'T','N' GEMM, M = nb*pc, N = nl, K = nb:
//           ri3ao viewed as the matrix [nb, nb*pc] needs no per-slab loop at all;
//       (2) for every a:  O[P, b] = sum_nu W[nu, P, a] C_R[nu,b]  -- strided-batched 'T','N' GEMM, M = pc (P!),
//           N = nr, K = nb, batch = nl, written straight into ri3mo[P + a*ldp + b*ldp*nl]: P is the M index of
//           the MMA tile, so the P-fastest output of the reference (a stride-naux scatter per element there)
//           becomes 128-byte contiguous stores.
//     The reference contracts nu first (dgemm NN then TN); here mu goes first.  Same sums, different rounding
//     order: agreement is ~1e-14 relative, the bar is 1e-10.
//
//   d_P, J, K  (not functions of the reference crate; composed as in SURVEY 3.5)
//       d_P = dgemv('T') over A = [nb^2, nx];  J = dgemv('N');  both HBM-bound, one pass over ri3ao each.
//       K   = sum_P (A_P Ct)(A_P Ct)^T : per chunk a strided-batched 'N','N' GEMM Y_P = A_P Ct followed by ONE
//             SYRK ('U','N') over the stacked Y = [nb, no*pc] (split-K, deterministic reduction), then mirrored.
#include "rb_common.cuh"

static i64 ws_budget_bytes(rb_ctx *ctx)
{
    // cudaMemGetInfo costs ~a millisecond: query once per context (workspaces are grow-only, so the first answer
    // stays a valid bound) instead of once per call.
    if (ctx->ws_budget > 0) return ctx->ws_budget;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); return (i64)1 << 30; }
    i64 have = (i64)free_b + ctx->ws_bytes[0];
    i64 budget = have / 3;
    // 16 GB by default: config D's whole per-rank intermediate (nb * 600 * nb doubles = 15.6 GB) then fits in ONE chunk,
    // which makes ao2mo GEMM 2 a single flat GEMM; REST_B200_WS_CAP_GB overrides.
    i64 cap = (i64)16 << 30;
    if (const char *e = getenv("REST_B200_WS_CAP_GB")) { const i64 v = atoll(e); if (v >= 1) cap = v << 30; }
    if (budget > cap) budget = cap;
    if (budget < ((i64)64 << 20)) budget = (i64)64 << 20;
    ctx->ws_budget = budget;
    return budget;
}

// Cost, in 128-row GEMM tiles, of a chunk of p slabs when P is the M index of the tile (ao2mo GEMM 2): full tiles plus
// the ragged one, which the kernel works on at 16-row block granularity with at most 2 block rows per warp row
// (1 block -> 1/8, 2 -> 1/4, 3-4 -> 1/2, 5-8 -> a full tile; see the warp-grid choice in rb_gemm.cu).
static double chunk_tile_cost(i64 p)
{
    const i64 full = p / 128, blocks = ((p % 128) + 15) / 16;
    const double rag = blocks == 0 ? 0.0 : blocks == 1 ? 0.125 : blocks == 2 ? 0.25 : blocks <= 4 ? 0.5 : 1.0;
    return (double)full + rag;
}

// P-chunk length: as large as the workspace budget allows.  tile_m: P is the M index of a GEMM tile -> among the chunk
// lengths that need the fewest chunks pick the one with the cheapest tiling (e.g. 600 slabs in two chunks: 320 + 280
// costs 4.75 tiles, 304 + 296 or 256 + 256 + 88 cost 5).
static i64 pick_chunk(i64 nx, i64 bytes_per_slab, i64 budget, bool tile_m)
{
    i64 pc_max = budget / (bytes_per_slab > 0 ? bytes_per_slab : 1);
    if (pc_max < 8) pc_max = 8;
    if (pc_max >= nx) return nx;
    const i64 nchunks = rb_cdiv(nx, pc_max);
    i64 pc = (rb_cdiv(nx, nchunks) + 7) & ~(i64)7;
    if (pc > pc_max) pc = pc_max & ~(i64)7;
    if (pc < 8) pc = 8;
    if (!tile_m) return pc < nx ? pc : nx;
    i64 best = pc;
    double best_cost = 1e300;
    for (i64 cand = pc; cand <= pc_max; cand += 8) {
        if (rb_cdiv(nx, cand) != nchunks) break;
        const i64 whole = nx / cand, rem = nx - whole * cand;
        const double cost = (double)whole * chunk_tile_cost(cand) + chunk_tile_cost(rem);
        if (cost < best_cost - 1e-9) { best_cost = cost; best = cand; }
    }
    return best < nx ? best : nx;
}

// host-only planner entry for the CPU tests (no device is touched)
extern "C" int64_t rb_ri_plan_chunk(int64_t nx, int64_t bytes_per_slab, int64_t budget_bytes, int tile_m)
{
    if (nx <= 0) return 0;
    return pick_chunk(nx, bytes_per_slab, budget_bytes, tile_m != 0);
}

// ---- odd nb / unaligned tensors ------------------------------------------------------------------------------
// TMA needs 16-byte aligned bases and row pitches, i.e. an even leading dimension.  Basis-set sizes are arbitrary, so
// for odd nb (or an 8-byte-aligned tensor) every P-chunk is first copied into zero-padded slabs [nbp, nbp], nbp = nb
// rounded up to even, C gets a zero row, and the same two GEMMs run with K = nbp: the extra products are exact zeros, so
// every sum keeps its value, at the cost of one HBM pass over the chunk (measured at nb = 601: the plain-load fallback
// GEMM kernel gives 5 TFLOP/s, this path the TMA kernel's rate minus the copy).
__global__ void __launch_bounds__(256) rb_pad_rows_kernel(const double *__restrict__ src, i64 n_in, i64 rows_in, i64 group_stride,
                                                          double *__restrict__ dst, i64 n_out, i64 rows_out, i64 total_rows)
{
    const int lane = threadIdx.x & 31;
    const i64 warps = (i64)gridDim.x * 8;
    for (i64 r = (i64)blockIdx.x * 8 + (threadIdx.x >> 5); r < total_rows; r += warps) {
        const i64 g = r / rows_out, v = r - g * rows_out;
        const bool valid = v < rows_in;
        const double *srow = src + g * group_stride + v * n_in;
        double *drow = dst + r * n_out;
        for (i64 i = lane; i < n_out; i += 32) drow[i] = (valid && i < n_in) ? srow[i] : 0.0;
    }
}

// dst[groups][rows_out][n_out] <- src[groups][rows_in][n_in] (dense), zero elsewhere
static int pad_rows(rb_ctx *ctx, const double *src, i64 n_in, i64 rows_in, double *dst, i64 n_out, i64 rows_out, i64 groups)
{
    const i64 total = rows_out * groups;
    if (total <= 0 || n_out <= 0) return RB_OK;
    i64 blocks = rb_cdiv(total, 8);
    const i64 cap = (i64)ctx->num_sms * 8;
    if (blocks > cap) blocks = cap;
    rb_pad_rows_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(src, n_in, rows_in, n_in * rows_in, dst, n_out, rows_out, total);
    RB_LAUNCHED(ctx);
    return RB_OK;
}

// dst[i + j*ldd] = beta*dst + src[i + j*lds] for the m x n block (upper: only i <= j)
__global__ void __launch_bounds__(256) rb_add_block_kernel(double *__restrict__ dst, i64 ldd, const double *__restrict__ src,
                                                           i64 lds, i64 m, i64 n, double beta, int upper)
{
    const i64 total = m * n, stride = (i64)gridDim.x * blockDim.x;
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const i64 i = e % m, j = e / m;
        if (upper && i > j) continue;
        double *d = dst + i + j * ldd;
        const double v = src[i + j * lds];
        *d = (beta == 0.0) ? v : beta * (*d) + v;
    }
}

static bool ri_needs_pad(i64 nb, const double *ri3ao) { return (nb & 1) || (((uintptr_t)ri3ao) & 15); }

static int ri_ao2mo_padded(rb_ctx *ctx, const double *c_left, i64 nl, const double *c_right, i64 nr, const double *ri3ao,
                           double *out, i64 nb, i64 nx, i64 out_ldp)
{
    const i64 nbp = nb + (nb & 1);
    const bool same_c = c_left == c_right && nl == nr;
    const i64 c_elems = nbp * nl + (same_c ? 0 : nbp * nr);
    i64 budget = ws_budget_bytes(ctx) - c_elems * 8;
    if (budget < ((i64)32 << 20)) budget = (i64)32 << 20;
    const i64 pc = pick_chunk(nx, (nbp * nbp + nbp * nl) * 8, budget, true);
    void *ws;
    RB_TRY(rb_ws_reserve(ctx, 0, (c_elems + nbp * nbp * pc + nbp * pc * nl) * 8, &ws));
    double *clp = (double *)ws, *crp = same_c ? clp : clp + nbp * nl;
    double *rip = clp + c_elems, *w = rip + nbp * nbp * pc;
    RB_TRY(pad_rows(ctx, c_left, nb, nl, clp, nbp, nl, 1));
    if (!same_c) RB_TRY(pad_rows(ctx, c_right, nb, nr, crp, nbp, nr, 1));
    for (i64 p0 = 0; p0 < nx; p0 += pc) {
        const i64 pn = (nx - p0 < pc) ? nx - p0 : pc;
        RB_TRY(pad_rows(ctx, ri3ao + p0 * nb * nb, nb, nb, rip, nbp, nbp, pn));
        RB_TRY(rb_gemm_core(ctx, true, false, nbp * pn, nl, nbp, 1.0, rip, nbp, 0, clp, nbp, 0, 0.0, w, nbp * pn, 0, 1, 0));
        if (pn == out_ldp)
            RB_TRY(rb_gemm_core(ctx, true, false, pn * nl, nr, nbp, 1.0, w, nbp, 0, crp, nbp, 0, 0.0, out, out_ldp * nl, 0, 1, 0));
        else
            RB_TRY(rb_gemm_core(ctx, true, false, pn, nr, nbp, 1.0, w, nbp, nbp * pn, crp, nbp, 0, 0.0, out + p0, out_ldp * nl,
                                out_ldp, nl, 0));
    }
    return RB_OK;
}

extern "C" int rb_ri_ao2mo(rb_ctx *ctx, const double *c_left, int nl, const double *c_right, int nr,
                           const double *ri3ao, double *out, int nb_, int nx_, int64_t out_ldp)
{
    RB_REQUIRE(ctx, "rb_ri_ao2mo: ctx is NULL");
    RB_REQUIRE(nl >= 0 && nr >= 0 && nb_ >= 0 && nx_ >= 0, "rb_ri_ao2mo: negative dimension");
    RB_REQUIRE(out_ldp >= nx_, "rb_ri_ao2mo: out_ldp (%lld) < nx (%d)", (long long)out_ldp, nx_);
    const i64 nb = nb_, nx = nx_;
    if (nx == 0 || nl == 0 || nr == 0) return RB_OK;
    RB_REQUIRE(out, "rb_ri_ao2mo: out is NULL");
    RB_CUDA(cudaSetDevice(ctx->device));
    if (nb == 0) { // empty contraction: ri3mo = 0 (the Fortran zero-fills first, restmatr.f90:178)
        for (i64 b = 0; b < nr; ++b)
            for (i64 a = 0; a < nl; ++a) RB_TRY(rb_scale_or_zero(ctx, out + a * out_ldp + b * out_ldp * nl, nx, 1, 0.0));
        return RB_OK;
    }
    RB_REQUIRE(c_left && c_right && ri3ao, "rb_ri_ao2mo: NULL input");
    if (ri_needs_pad(nb, ri3ao) || (((uintptr_t)c_left | (uintptr_t)c_right) & 15))
        return ri_ao2mo_padded(ctx, c_left, nl, c_right, nr, ri3ao, out, nb, nx, out_ldp);
    i64 pc = pick_chunk(nx, nb * nl * 8, ws_budget_bytes(ctx), true);
    void *ws;
    int st = rb_ws_reserve(ctx, 0, nb * pc * nl * 8, &ws);
    if (st == RB_ERR_NOMEM) { // the budget was stale (rb_ws_reserve reset it): plan again against what is free now
        pc = pick_chunk(nx, nb * nl * 8, ws_budget_bytes(ctx), true);
        st = rb_ws_reserve(ctx, 0, nb * pc * nl * 8, &ws);
    }
    RB_TRY(st);
    double *w = (double *)ws;
    for (i64 p0 = 0; p0 < nx; p0 += pc) {
        const i64 pn = (nx - p0 < pc) ? nx - p0 : pc;
        // (1) W[(nu,P), a] : 'T','N'  M = nb*pn, N = nl, K = nb
        RB_TRY(rb_gemm_core(ctx, true, false, nb * pn, nl, nb, 1.0, ri3ao + p0 * nb * nb, nb, 0, c_left, nb, 0, 0.0, w,
                            nb * pn, 0, 1, 0));
        // (2) per a: O_a[P, b] : 'T','N'  M = pn, N = nr, K = nb ; A = W_a [nb x pn], C = out + p0 + a*ldp, ldc = ldp*nl.
        // When the chunk is the whole P range of `out` (pn == ldp) the rows (P, a) of all a are one contiguous run in
        // both W and out, so the nl batches collapse into ONE GEMM with M = pn*nl: no ragged tile row per a.
        if (pn == out_ldp)
            RB_TRY(rb_gemm_core(ctx, true, false, pn * nl, nr, nb, 1.0, w, nb, 0, c_right, nb, 0, 0.0, out, out_ldp * nl, 0,
                                1, 0));
        else
            RB_TRY(rb_gemm_core(ctx, true, false, pn, nr, nb, 1.0, w, nb, nb * pn, c_right, nb, 0, 0.0, out + p0,
                                out_ldp * nl, out_ldp, nl, 0));
    }
    return RB_OK;
}

extern "C" int rb_ri_dp(rb_ctx *ctx, const double *ri3ao, const double *dm, double *d, int nb, int nx)
{
    RB_REQUIRE(ctx, "rb_ri_dp: ctx is NULL");
    RB_REQUIRE(nb >= 0 && nx >= 0, "rb_ri_dp: negative dimension");
    if (nx == 0) return RB_OK;
    RB_CUDA(cudaSetDevice(ctx->device));
    const i64 m = (i64)nb * nb;
    if (m == 0) return rb_scale_or_zero(ctx, d, nx, 1, 0.0);
    RB_REQUIRE(m <= 2147483647LL, "rb_ri_dp: nb too large");
    return rb_dgemv(ctx, 'T', (int)m, nx, 1.0, ri3ao, m, dm, 1, 0.0, d, 1);
}

extern "C" int rb_ri_j(rb_ctx *ctx, const double *ri3ao, const double *d, double *j, int nb, int nx)
{
    RB_REQUIRE(ctx, "rb_ri_j: ctx is NULL");
    RB_REQUIRE(nb >= 0 && nx >= 0, "rb_ri_j: negative dimension");
    const i64 m = (i64)nb * nb;
    if (m == 0) return RB_OK;
    RB_CUDA(cudaSetDevice(ctx->device));
    if (nx == 0) return rb_scale_or_zero(ctx, j, m, 1, 0.0);
    RB_REQUIRE(m <= 2147483647LL, "rb_ri_j: nb too large");
    return rb_dgemv(ctx, 'N', (int)m, nx, 1.0, ri3ao, m, d, 1, 0.0, j, 1);
}

// Upper triangle of k (+)= sum_P Y_P Y_P^T over the given slabs (beta = 0 overwrites, 1 accumulates); no mirroring.
int rb_ri_k_upper(rb_ctx *ctx, const double *ri3ao, const double *ct, i64 no, double *k, i64 nb, i64 nx, double beta)
{
    if (ri_needs_pad(nb, ri3ao) || (((uintptr_t)ct) & 15)) { // see the note above ri_ao2mo_padded
        const i64 nbp = nb + (nb & 1);
        const i64 fixed = nbp * no + nbp * nbp; // padded Ct and the padded accumulator K'
        i64 budget = ws_budget_bytes(ctx) - fixed * 8;
        if (budget < ((i64)32 << 20)) budget = (i64)32 << 20;
        const i64 pc = pick_chunk(nx, (nbp * nbp + nbp * no) * 8, budget, false);
        RB_REQUIRE(no * pc <= 2147483647LL, "rb_ri_k: chunk too large");
        void *ws;
        RB_TRY(rb_ws_reserve(ctx, 0, (fixed + nbp * nbp * pc + nbp * no * pc) * 8, &ws));
        double *ctp = (double *)ws, *kp = ctp + nbp * no, *rip = kp + nbp * nbp, *y = rip + nbp * nbp * pc;
        RB_TRY(pad_rows(ctx, ct, nb, no, ctp, nbp, no, 1));
        for (i64 p0 = 0; p0 < nx; p0 += pc) {
            const i64 pn = (nx - p0 < pc) ? nx - p0 : pc;
            RB_TRY(pad_rows(ctx, ri3ao + p0 * nb * nb, nb, nb, rip, nbp, nbp, pn));
            RB_TRY(rb_gemm_core(ctx, false, false, nbp, no, nbp, 1.0, rip, nbp, nbp * nbp, ctp, nbp, 0, 0.0, y, nbp, nbp * no, pn, 0));
            RB_TRY(rb_gemm_core(ctx, false, true, nbp, nbp, no * pn, 1.0, y, nbp, 0, y, nbp, 0, p0 == 0 ? 0.0 : 1.0, kp, nbp, 0, 1, 1));
        }
        i64 blocks = rb_cdiv(nb * nb, 256);
        if (blocks > (i64)ctx->num_sms * 8) blocks = (i64)ctx->num_sms * 8;
        rb_add_block_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(k, nb, kp, nbp, nb, nb, beta, 1);
        RB_LAUNCHED(ctx);
        return RB_OK;
    }
    i64 pc = pick_chunk(nx, nb * no * 8, ws_budget_bytes(ctx), false);
    RB_REQUIRE(no * pc <= 2147483647LL, "rb_ri_k: chunk too large");
    void *ws;
    int st = rb_ws_reserve(ctx, 0, nb * no * pc * 8, &ws);
    if (st == RB_ERR_NOMEM) { // stale budget: plan again against what is free now
        pc = pick_chunk(nx, nb * no * 8, ws_budget_bytes(ctx), false);
        st = rb_ws_reserve(ctx, 0, nb * no * pc * 8, &ws);
    }
    RB_TRY(st);
    double *y = (double *)ws;
    for (i64 p0 = 0; p0 < nx; p0 += pc) {
        const i64 pn = (nx - p0 < pc) ? nx - p0 : pc;
        // Y_P = A_P * Ct : 'N','N'  M = nb, N = no, K = nb, batch = pn
        RB_TRY(rb_gemm_core(ctx, false, false, nb, no, nb, 1.0, ri3ao + p0 * nb * nb, nb, nb * nb, ct, nb, 0, 0.0, y, nb,
                            nb * no, pn, 0));
        // K(upper) (+)= Y Y^T : SYRK 'U','N' with k = no*pn
        RB_TRY(rb_gemm_core(ctx, false, true, nb, nb, no * pn, 1.0, y, nb, 0, y, nb, 0, p0 == 0 ? beta : 1.0, k, nb, 0, 1,
                            1));
    }
    return RB_OK;
}

extern "C" int rb_ri_k(rb_ctx *ctx, const double *ri3ao, const double *ct, int no_, double *k, int nb_, int nx_)
{
    RB_REQUIRE(ctx, "rb_ri_k: ctx is NULL");
    RB_REQUIRE(no_ >= 0 && nb_ >= 0 && nx_ >= 0, "rb_ri_k: negative dimension");
    const i64 nb = nb_, nx = nx_, no = no_;
    if (nb == 0) return RB_OK;
    RB_REQUIRE(k, "rb_ri_k: k is NULL");
    RB_CUDA(cudaSetDevice(ctx->device));
    if (nx == 0 || no == 0) return rb_scale_or_zero(ctx, k, nb * nb, 1, 0.0);
    RB_REQUIRE(ct && ri3ao, "rb_ri_k: NULL input");
    RB_TRY(rb_ri_k_upper(ctx, ri3ao, ct, no, k, nb, nx, 0.0));
    return rb_symmetrize(ctx, k, nb, nb, true);
}

// restmatr.f90:111-154 on device buffers: for every y, T[xr, y, zr] <- alpha * T[xr, y, zr] * B[:, 0..len_col_b) + beta * T[..]
// (in place => len_col_b must equal len_z).  b points at the first element of the B block (ldb = rows of B).
extern "C" int rb_special_dgemm_01(rb_ctx *ctx, double *ten3, int x_a, int y_a, int z_a, int start_x, int len_x,
                                   int start_z, int len_z, const double *b, int64_t ldb, int len_col_b, double alpha,
                                   double beta)
{
    RB_REQUIRE(ctx, "rb_special_dgemm_01: ctx is NULL");
    RB_REQUIRE(x_a >= 0 && y_a >= 0 && z_a >= 0 && start_x >= 0 && len_x >= 0 && start_z >= 0 && len_z >= 0,
               "rb_special_dgemm_01: negative dimension");
    RB_REQUIRE(start_x + len_x <= x_a && start_z + len_z <= z_a, "rb_special_dgemm_01: block outside tensor");
    RB_REQUIRE(len_col_b == len_z, "rb_special_dgemm_01: in-place update needs len_column_b == len_z_a (%d vs %d)",
               len_col_b, len_z);
    if (len_x == 0 || len_z == 0 || y_a == 0) return RB_OK;
    RB_CUDA(cudaSetDevice(ctx->device));
    const i64 X = x_a, Y = y_a;
    const i64 lx = len_x, lz = len_z;
    // process y in chunks bounded by the workspace budget
    i64 yc = ws_budget_bytes(ctx) / (lx * lz * 8);
    if (yc < 1) yc = 1;
    if (yc > Y) yc = Y;
    void *ws;
    RB_TRY(rb_ws_reserve(ctx, 0, lx * lz * yc * 8, &ws));
    double *tmp = (double *)ws;
    for (i64 y0 = 0; y0 < Y; y0 += yc) {
        const i64 yn = (Y - y0 < yc) ? Y - y0 : yc;
        double *blk = ten3 + start_x + y0 * X + (i64)start_z * X * Y;
        // tmp[x, z, y] = T[sx+x, y0+y, sz+z]
        RB_TRY(rb_copy3d(ctx, blk, 0, 1, X * Y, X, tmp, 0, 1, lx, lx * lz, lx, lz, yn));
        // T block (ldc = X*Y, batch stride X) = alpha * tmp_y * B + beta * T block
        RB_TRY(rb_gemm_core(ctx, false, false, lx, lz, lz, alpha, tmp, lx, lx * lz, b, ldb, 0, beta, blk, X * Y, X, yn, 0));
    }
    return RB_OK;
}

// ---- (ia|jb)-type consumers of ri3mo -------------------------------------------------------
// ri3mo is P-fastest, mo[P + l*ldp + r*ldp*nl] (reference src/ri.rs:381-386): a box of MO pairs (l in [l0,l0+ll),
// r in [r0,r0+rl)) is a set of length-np columns, so
//     out[(l,r)_A, (l',r')_B] (+)= sum_P moA[P,l,r] * moB[P,l',r']
// is ONE 'T','N' DMMA GEMM with K = np and both operands K-major (the layout ao2mo writes is already the layout this
// GEMM reads; no transposition anywhere).  A box that spans the whole l range is a contiguous column block of mo and
// is read in place; a partial l range is gathered once into a dense workspace panel (one HBM pass, <= 1/M of the
// GEMM's work).  Same tensor and same box on both sides: SYRK (upper-triangle tiles only) + mirror.  moA != moB covers
// the alpha/beta spin blocks.  P-sharded ranks each produce a partial sum over their local rows; the caller all-reduces.
struct MoBox { const double *mo; i64 ldp, nl, nr, l0, ll, r0, rl; };

static bool box_is_panel(const MoBox &x) { return x.l0 == 0 && x.ll == x.nl; }
static bool box_valid(const MoBox &x)
{
    return x.nl >= 0 && x.nr >= 0 && x.l0 >= 0 && x.ll >= 0 && x.l0 + x.ll <= x.nl && x.r0 >= 0 && x.rl >= 0 && x.r0 + x.rl <= x.nr;
}

extern "C" int rb_ri_iajb(rb_ctx *ctx, int np_, const double *mo_a, int64_t ldp_a, int nl_a, int nr_a, int l0a, int lla,
                          int r0a, int rla, const double *mo_b, int64_t ldp_b, int nl_b, int nr_b, int l0b, int llb,
                          int r0b, int rlb, double beta, double *out, int64_t ldo)
{
    RB_REQUIRE(ctx, "rb_ri_iajb: ctx is NULL");
    RB_REQUIRE(np_ >= 0, "rb_ri_iajb: negative dimension");
    const MoBox A = {mo_a, ldp_a, nl_a, nr_a, l0a, lla, r0a, rla}, B = {mo_b, ldp_b, nl_b, nr_b, l0b, llb, r0b, rlb};
    RB_REQUIRE(box_valid(A), "rb_ri_iajb: box A [%d+%d, %d+%d] outside [%d, %d]", l0a, lla, r0a, rla, nl_a, nr_a);
    RB_REQUIRE(box_valid(B), "rb_ri_iajb: box B [%d+%d, %d+%d] outside [%d, %d]", l0b, llb, r0b, rlb, nl_b, nr_b);
    RB_REQUIRE(ldp_a >= np_ && ldp_b >= np_, "rb_ri_iajb: ldp (%lld, %lld) < np (%d)", (long long)ldp_a, (long long)ldp_b, np_);
    const i64 np = np_, m = A.ll * A.rl, n = B.ll * B.rl;
    if (m == 0 || n == 0) return RB_OK;
    RB_REQUIRE(out, "rb_ri_iajb: out is NULL");
    RB_REQUIRE(ldo >= m, "rb_ri_iajb: ldo (%lld) < rows of the block (%lld)", (long long)ldo, (long long)m);
    RB_CUDA(cudaSetDevice(ctx->device));
    if (np == 0) // empty contraction: out = beta * out
        return rb_gemm_core(ctx, true, false, m, n, 0, 1.0, nullptr, 1, 0, nullptr, 1, 0, beta, out, ldo, 0, 1, 0);
    RB_REQUIRE(mo_a && mo_b, "rb_ri_iajb: mo is NULL");
    const double *xa = mo_a + A.l0 * A.ldp + A.r0 * A.ldp * A.nl, *xb = mo_b + B.l0 * B.ldp + B.r0 * B.ldp * B.nl;
    const bool same = xa == xb && A.ldp == B.ldp && A.nl == B.nl && A.ll == B.ll && A.rl == B.rl;
    // SYRK form (upper-triangle tiles + mirror) only when out is overwritten: with beta != 0 the caller's out need not be
    // symmetric, so the full product is formed like any GEMM
    const int tri = (same && beta == 0.0) ? 1 : 0;
    const bool pa = box_is_panel(A), pb = same ? pa : box_is_panel(B);
    if (pa && pb) { // both boxes are column panels of mo: read in place
        RB_TRY(rb_gemm_core(ctx, true, false, m, n, np, 1.0, xa, A.ldp, 0, xb, B.ldp, 0, beta, out, ldo, 0, 1, tri));
        return tri ? rb_symmetrize(ctx, out, m, ldo, true) : RB_OK;
    }
    // gather the non-panel boxes, P-chunked so the panels fit the workspace budget (chunks accumulate with beta = 1)
    const i64 cols = (pa ? 0 : m) + ((pb || same) ? 0 : n);
    i64 pc = (ws_budget_bytes(ctx) / (cols * 8)) & ~(i64)7;
    if (pc < 8) pc = 8;
    if (pc > np) pc = np;
    const i64 ldx = pc + (pc & 1); // even pitch: the TMA path needs 16-byte aligned columns
    void *ws;
    RB_TRY(rb_ws_reserve(ctx, 0, ldx * cols * 8, &ws));
    double *ga = (double *)ws, *gb = pa ? ga : ga + ldx * m;
    for (i64 p0 = 0; p0 < np; p0 += pc) {
        const i64 pn = (np - p0 < pc) ? np - p0 : pc;
        const double *a = xa + p0, *b = xb + p0;
        i64 lda = A.ldp, ldb = B.ldp;
        if (!pa) {
            RB_TRY(rb_copy3d(ctx, xa + p0, 0, 1, A.ldp, A.ldp * A.nl, ga, 0, 1, ldx, ldx * A.ll, pn, A.ll, A.rl));
            a = ga; lda = ldx;
        }
        if (same) { b = a; ldb = lda; }
        else if (!pb) {
            RB_TRY(rb_copy3d(ctx, xb + p0, 0, 1, B.ldp, B.ldp * B.nl, gb, 0, 1, ldx, ldx * B.ll, pn, B.ll, B.rl));
            b = gb; ldb = ldx;
        }
        RB_TRY(rb_gemm_core(ctx, true, false, m, n, pn, 1.0, a, lda, 0, b, ldb, 0, p0 == 0 ? beta : 1.0, out, ldo, 0, 1, tri));
    }
    return tri ? rb_symmetrize(ctx, out, m, ldo, true) : RB_OK;
}

// ---- RPA-type consumer of ri3mo: contraction over the MO pairs, output in the auxiliary basis ------------------------
//     out[P, Q] (+)= sum_{(l,r) in box} w[l,r] * moA[P,l,r] * moB[Q,l,r]        (w == NULL: all ones)
// e.g. the RPA polarisability Pi(omega) = sum_ia R_ia^P f_ia(omega) R_ia^Q, which REST builds with _dgemm on
// [naux, n_occ*n_vir] views of ri3mo.  moA / moB are row blocks (P-shards) of the same P-fastest tensor: a rank that
// owns rows P_r gets out[P_r, Q_s] from its own rows and rank s's rows.  ONE 'N','T' DMMA GEMM with K = the box's MO
// pairs; both operands are MN-major (P fastest), i.e. the swizzled TMA path.  The weights are folded into a scaled
// copy of the B panel (one HBM pass); a partial l range is gathered by the same kernel.  moA == moB: only the upper
// triangle's tiles are computed, then mirrored (sum_c w_c x_c x_c^T is symmetric).
__global__ void __launch_bounds__(256) rb_mo_box_gather_kernel(const double *__restrict__ src, i64 ldp, i64 nl, i64 ll,
                                                               i64 c0, const double *__restrict__ w,
                                                               double *__restrict__ dst, i64 ldd, i64 np, i64 cols)
{
    for (i64 c = blockIdx.y; c < cols; c += gridDim.y) {
        const i64 cb = c0 + c, l = cb % ll, r = cb / ll;
        const double *s = src + l * ldp + r * ldp * nl;
        double *d = dst + c * ldd;
        const i64 stride = (i64)gridDim.x * blockDim.x;
        i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x;
        if (w) {
            const double wc = w[cb];
            for (; p + stride < np; p += 2 * stride) {
                const double v0 = s[p], v1 = s[p + stride];
                d[p] = __dmul_rn(v0, wc); d[p + stride] = __dmul_rn(v1, wc);
            }
            for (; p < np; p += stride) d[p] = __dmul_rn(s[p], wc);
        } else {
            for (; p + stride < np; p += 2 * stride) {
                const double v0 = s[p], v1 = s[p + stride];
                d[p] = v0; d[p + stride] = v1;
            }
            for (; p < np; p += stride) d[p] = s[p];
        }
    }
}

static int mo_box_gather(rb_ctx *ctx, const double *box, i64 ldp, i64 nl, i64 ll, i64 c0, const double *w, double *dst,
                         i64 ldd, i64 np, i64 cols)
{
    i64 bx = rb_cdiv(np, 512);
    if (bx > 16) bx = 16;
    i64 by = cols < 65535 ? cols : 65535;
    const i64 cap = (i64)ctx->num_sms * 16;
    if (bx * by > cap) by = cap / bx > 0 ? cap / bx : 1;
    rb_mo_box_gather_kernel<<<dim3((unsigned)bx, (unsigned)by), 256, 0, ctx->stream>>>(box, ldp, nl, ll, c0, w, dst, ldd, np, cols);
    RB_LAUNCHED(ctx);
    return RB_OK;
}

extern "C" int rb_ri_mo_pq(rb_ctx *ctx, const double *mo_a, int64_t ldp_a, int np_a, const double *mo_b, int64_t ldp_b,
                           int np_b, int nl_, int nr_, int l0, int ll_, int r0, int rl_, const double *w, double beta,
                           double *out, int64_t ldo)
{
    RB_REQUIRE(ctx, "rb_ri_mo_pq: ctx is NULL");
    RB_REQUIRE(np_a >= 0 && np_b >= 0, "rb_ri_mo_pq: negative dimension");
    const MoBox X = {mo_a, ldp_a, nl_, nr_, l0, ll_, r0, rl_};
    RB_REQUIRE(box_valid(X), "rb_ri_mo_pq: box [%d+%d, %d+%d] outside [%d, %d]", l0, ll_, r0, rl_, nl_, nr_);
    RB_REQUIRE(ldp_a >= np_a && ldp_b >= np_b, "rb_ri_mo_pq: ldp (%lld, %lld) < np (%d, %d)", (long long)ldp_a,
               (long long)ldp_b, np_a, np_b);
    const i64 m = np_a, n = np_b, nl = nl_, ll = ll_, cols = (i64)ll_ * rl_;
    if (m == 0 || n == 0) return RB_OK;
    RB_REQUIRE(out, "rb_ri_mo_pq: out is NULL");
    RB_REQUIRE(ldo >= m, "rb_ri_mo_pq: ldo (%lld) < np_a (%lld)", (long long)ldo, (long long)m);
    RB_CUDA(cudaSetDevice(ctx->device));
    if (cols == 0) // empty contraction: out = beta * out
        return rb_gemm_core(ctx, false, true, m, n, 0, 1.0, nullptr, 1, 0, nullptr, 1, 0, beta, out, ldo, 0, 1, 0);
    RB_REQUIRE(mo_a && mo_b, "rb_ri_mo_pq: mo is NULL");
    const bool same = mo_a == mo_b && ldp_a == ldp_b && np_a == np_b;
    const int tri = (same && beta == 0.0) ? 1 : 0; // beta != 0: out need not be symmetric on entry -> full product
    const bool panel = box_is_panel(X);
    const double *xa = mo_a + (i64)l0 * ldp_a + (i64)r0 * ldp_a * nl, *xb = mo_b + (i64)l0 * ldp_b + (i64)r0 * ldp_b * nl;
    const bool copy_a = !panel, copy_b = !panel || w != nullptr;
    if (!copy_a && !copy_b) {
        RB_TRY(rb_gemm_core(ctx, false, true, m, n, cols, 1.0, xa, ldp_a, 0, xb, ldp_b, 0, beta, out, ldo, 0, 1, tri));
        return tri ? rb_symmetrize(ctx, out, m, ldo, true) : RB_OK;
    }
    // column-chunked: gathered / weighted copies of the panels live in the workspace, chunks accumulate with beta = 1
    const i64 lda_g = m + (m & 1), ldb_g = n + (n & 1);
    const i64 per_col = ((copy_a ? lda_g : 0) + (copy_b ? ldb_g : 0)) * 8;
    i64 cc = ws_budget_bytes(ctx) / per_col;
    if (cc > 8) cc &= ~(i64)7;
    if (cc < 1) cc = 1;
    if (cc > cols) cc = cols;
    void *ws;
    RB_TRY(rb_ws_reserve(ctx, 0, per_col * cc, &ws));
    double *ga = (double *)ws, *gb = copy_a ? ga + lda_g * cc : ga;
    for (i64 c0 = 0; c0 < cols; c0 += cc) {
        const i64 cn = (cols - c0 < cc) ? cols - c0 : cc;
        const double *a = xa + c0 * ldp_a, *b = gb;
        i64 lda = ldp_a;
        if (copy_a) {
            RB_TRY(mo_box_gather(ctx, xa, ldp_a, nl, ll, c0, nullptr, ga, lda_g, m, cn));
            a = ga; lda = lda_g;
        }
        RB_TRY(mo_box_gather(ctx, xb, ldp_b, nl, ll, c0, w, gb, ldb_g, n, cn));
        RB_TRY(rb_gemm_core(ctx, false, true, m, n, cn, 1.0, a, lda, 0, b, ldb_g, 0, c0 == 0 ? beta : 1.0, out, ldo, 0, 1, tri));
    }
    return tri ? rb_symmetrize(ctx, out, m, ldo, true) : RB_OK;
}

// ---- all-gather -> GEMM as one pipeline over peer memory ---------------------------------------------------------------
// P-sharded form of the RPA-type consumer: rank r owns rows P_r of the box panel X = [naux, cols] and needs
//     out[P_r, Q_s] = sum_c w_c X[P_r, c] X[Q_s, c]      for every rank s.
// panels[s] is rank s's dense panel [np[s] rows, pitch ld, cols columns] as seen from this rank (a CUDA IPC mapping of
// the peer's HBM, or a local pointer for s == rank).  Measured on 2 x B200 (tools/peer_probe.py): a GEMM whose TMA loads
// read the peer panel in place re-reads it once per tile row over NVLink (850 rows: 10.7 TFLOP/s vs 32.1 local), so the
// panels are pulled instead -- by the copy engines, on a second stream, one peer ahead of the DMMA GEMM that consumes
// them (double-buffered): NVLink transfer of peer s+1 overlaps the math on peer s, the own block is computed while the
// first pull is in flight, and no SM time goes into communication.  The weights are folded into a scaled copy of the
// OWN panel (A side), so the pulled panels are used as they are.
extern "C" int rb_ri_mo_pq_peers(rb_ctx *ctx, int rank, int world, const double *const *panels, int64_t ld, const int *np,
                                 int64_t cols, const double *w, double *out, int64_t ldo, const int64_t *q_off)
{
    RB_REQUIRE(ctx && panels && np && q_off, "rb_ri_mo_pq_peers: NULL argument");
    RB_NO_CAPTURE(ctx, "rb_ri_mo_pq_peers");
    RB_REQUIRE(world >= 1 && rank >= 0 && rank < world && cols >= 0, "rb_ri_mo_pq_peers: bad rank / world / cols");
    const i64 m = np[rank];
    RB_REQUIRE(m >= 0 && ldo >= m, "rb_ri_mo_pq_peers: ldo (%lld) < local rows (%lld)", (long long)ldo, (long long)m);
    i64 np_max = 0;
    for (int s = 0; s < world; ++s) {
        RB_REQUIRE(np[s] >= 0 && np[s] <= ld && q_off[s] >= 0, "rb_ri_mo_pq_peers: bad row count / offset of rank %d", s);
        RB_REQUIRE(np[s] == 0 || cols == 0 || panels[s], "rb_ri_mo_pq_peers: panel of rank %d is NULL", s);
        if (np[s] > np_max) np_max = np[s];
    }
    if (m == 0) return RB_OK;
    RB_REQUIRE(out, "rb_ri_mo_pq_peers: out is NULL");
    RB_CUDA(cudaSetDevice(ctx->device));
    if (cols == 0) {
        for (int s = 0; s < world; ++s)
            RB_TRY(rb_gemm_core(ctx, false, true, m, np[s], 0, 1.0, nullptr, 1, 0, nullptr, 1, 0, 0.0, out + q_off[s] * ldo, ldo, 0, 1, 0));
        return RB_OK;
    }
    if (!ctx->aux_stream) {
        RB_CUDA(cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 5; ++i) RB_CUDA(cudaEventCreateWithFlags(&ctx->aux_ev[i], cudaEventDisableTiming));
    }
    const i64 panel_elems = ld * cols;
    const int remote = world - 1;
    void *ws;
    RB_TRY(rb_ws_reserve(ctx, 0, (panel_elems * ((w ? 1 : 0) + (remote > 1 ? 2 : remote))) * 8 + 16, &ws));
    double *aw = (double *)ws;                                    // own panel * diag(w)
    double *buf[2] = {aw + (w ? panel_elems : 0), aw + (w ? panel_elems : 0) + panel_elems};
    const double *a = panels[rank];
    // the copy stream starts after everything already queued on the main stream (workspace reuse by earlier calls)
    RB_CUDA(cudaEventRecord(ctx->aux_ev[0], ctx->stream));
    RB_CUDA(cudaStreamWaitEvent(ctx->aux_stream, ctx->aux_ev[0], 0));
    auto peer_of = [&](int i) { return (rank + 1 + i) % world; };  // ring order: spreads the NVLink load over the switch
    auto pull = [&](int i) -> int {
        const int s = peer_of(i), b = i & 1;
        if (i >= 2) RB_CUDA(cudaStreamWaitEvent(ctx->aux_stream, ctx->aux_ev[3 + b], 0)); // the GEMM that read buf[b] is done
        if (np[s] > 0) // the whole panel as one run (its <= 2 pad rows per column travel too and are never read)
            RB_CUDA(cudaMemcpyAsync(buf[b], panels[s], (size_t)panel_elems * 8, cudaMemcpyDefault, ctx->aux_stream));
        RB_CUDA(cudaEventRecord(ctx->aux_ev[1 + b], ctx->aux_stream));
        return RB_OK;
    };
    if (remote > 0) RB_TRY(pull(0));
    if (w) { RB_TRY(mo_box_gather(ctx, a, ld, 1, 1, 0, w, aw, ld, m, cols)); a = aw; }
    // own block while the first pull is in flight: sum_c w_c x_c x_c^T is symmetric with or without weights, so only the
    // upper triangle's tiles are computed and mirrored
    RB_TRY(rb_gemm_core(ctx, false, true, m, m, cols, 1.0, a, ld, 0, panels[rank], ld, 0, 0.0, out + q_off[rank] * ldo, ldo, 0, 1, 1));
    RB_TRY(rb_symmetrize(ctx, out + q_off[rank] * ldo, m, ldo, true));
    for (int i = 0; i < remote; ++i) {
        const int s = peer_of(i), b = i & 1;
        if (i + 1 < remote) RB_TRY(pull(i + 1));
        RB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->aux_ev[1 + b], 0));
        if (np[s] > 0)
            RB_TRY(rb_gemm_core(ctx, false, true, m, np[s], cols, 1.0, a, ld, 0, buf[b], ld, 0, 0.0, out + q_off[s] * ldo, ldo, 0, 1, 0));
        RB_CUDA(cudaEventRecord(ctx->aux_ev[3 + b], ctx->stream));
    }
    return RB_OK;
}

// ---- special_dgemm_f_01 contracted over the SHARDED index -------------------------------------------
// With the full x range the reference's per-y loop T[:, y, :] <- alpha T[:, y, :] B + beta T[:, y, :] is ONE matrix product
// on the [X*Y, naux] view of the tensor, and with P-sharding the contraction index is the sharded one: rank s needs
//     out_s[xy, P'_s] = alpha * sum_r T_r[xy, P_r] B[P_r, P'_s] + beta * T_s[xy, P'_s].
// Same pipeline as rb_ri_mo_pq_peers, at row-chunk granularity (a whole shard is up to 15.5 GB at config D): the copy
// engines pull chunk (c, r+1) of peer r+1's shard -- a pitched [rows, np_r] block -- while the 'N','N' DMMA GEMM on chunk
// (c, r) accumulates into out; the own shard's term needs no pull and opens every chunk (beta * T_s is folded into it).
extern "C" int rb_special_dgemm_01_peers(rb_ctx *ctx, int rank, int world, const double *const *shards, int64_t xy,
                                         const int *np, const int64_t *p_off, const double *b, int64_t ldb, double alpha,
                                         double beta, double *out)
{
    RB_REQUIRE(ctx && shards && np && p_off, "rb_special_dgemm_01_peers: NULL argument");
    RB_NO_CAPTURE(ctx, "rb_special_dgemm_01_peers");
    RB_REQUIRE(world >= 1 && rank >= 0 && rank < world && xy >= 0, "rb_special_dgemm_01_peers: bad rank / world / rows");
    i64 naux = 0, np_max = 0;
    for (int s = 0; s < world; ++s) {
        RB_REQUIRE(np[s] >= 0 && p_off[s] >= 0, "rb_special_dgemm_01_peers: bad shard of rank %d", s);
        RB_REQUIRE(np[s] == 0 || xy == 0 || shards[s], "rb_special_dgemm_01_peers: shard of rank %d is NULL", s);
        if (p_off[s] + np[s] > naux) naux = p_off[s] + np[s];
        if (s != rank && np[s] > np_max) np_max = np[s];
    }
    const i64 n = np[rank];
    if (n == 0 || xy == 0) return RB_OK;
    RB_REQUIRE(out && b, "rb_special_dgemm_01_peers: NULL buffer");
    RB_REQUIRE(ldb >= naux, "rb_special_dgemm_01_peers: ldb (%lld) < naux (%lld)", (long long)ldb, (long long)naux);
    RB_CUDA(cudaSetDevice(ctx->device));
    if (!ctx->aux_stream) {
        RB_CUDA(cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 5; ++i) RB_CUDA(cudaEventCreateWithFlags(&ctx->aux_ev[i], cudaEventDisableTiming));
    }
    const int remote = world - 1;
    // row chunk: two pull buffers of [mc, np_max] within a quarter of the workspace budget; multiples of 128 rows (GEMM tiles)
    i64 mc = xy;
    if (remote > 0 && np_max > 0) {
        mc = (ws_budget_bytes(ctx) / 4) / (2 * np_max * 8);
        mc &= ~(i64)127;
        if (mc < 128) mc = 128;
        if (mc > xy) mc = xy;
    }
    const i64 ldbuf = mc + (mc & 1);
    double *buf[2] = {nullptr, nullptr};
    if (remote > 0 && np_max > 0) {
        void *ws;
        RB_TRY(rb_ws_reserve(ctx, 0, 2 * ldbuf * np_max * 8, &ws));
        buf[0] = (double *)ws; buf[1] = buf[0] + ldbuf * np_max;
    }
    RB_CUDA(cudaEventRecord(ctx->aux_ev[0], ctx->stream));
    RB_CUDA(cudaStreamWaitEvent(ctx->aux_stream, ctx->aux_ev[0], 0));
    const i64 nchunks = rb_cdiv(xy, mc), steps = nchunks * remote;
    auto step_of = [&](i64 t, i64 &row0, i64 &rows, int &peer) {
        const i64 c = t / remote;
        row0 = c * mc; rows = (xy - row0 < mc) ? xy - row0 : mc;
        peer = (rank + 1 + (int)(t % remote)) % world;
    };
    auto pull = [&](i64 t) -> int {
        i64 row0, rows; int r;
        step_of(t, row0, rows, r);
        const int bsel = (int)(t & 1);
        if (t >= 2) RB_CUDA(cudaStreamWaitEvent(ctx->aux_stream, ctx->aux_ev[3 + bsel], 0));
        if (np[r] > 0)
            RB_CUDA(cudaMemcpy2DAsync(buf[bsel], (size_t)ldbuf * 8, shards[r] + row0, (size_t)xy * 8, (size_t)rows * 8, (size_t)np[r],
                                      cudaMemcpyDefault, ctx->aux_stream));
        RB_CUDA(cudaEventRecord(ctx->aux_ev[1 + bsel], ctx->aux_stream));
        return RB_OK;
    };
    if (steps > 0) RB_TRY(pull(0));
    const double *bcol = b + p_off[rank] * ldb; // columns P'_s of B
    i64 t = 0;
    for (i64 c = 0; c < nchunks; ++c) {
        const i64 row0 = c * mc, rows = (xy - row0 < mc) ? xy - row0 : mc;
        double *oc = out + row0;
        // own term: out = alpha * T_s B[P_s, P'_s] + beta * T_s   (out starts as a copy of T_s when beta != 0)
        if (beta != 0.0) RB_TRY(rb_copy3d(ctx, shards[rank] + row0, 0, 1, xy, 0, oc, 0, 1, xy, 0, rows, n, 1));
        RB_TRY(rb_gemm_core(ctx, false, false, rows, n, n, alpha, shards[rank] + row0, xy, 0, bcol + p_off[rank], ldb, 0, beta, oc, xy, 0, 1, 0));
        for (int i = 0; i < remote; ++i, ++t) {
            i64 r0, rr; int r;
            step_of(t, r0, rr, r);
            const int bsel = (int)(t & 1);
            if (t + 1 < steps) RB_TRY(pull(t + 1));
            RB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->aux_ev[1 + bsel], 0));
            if (np[r] > 0)
                RB_TRY(rb_gemm_core(ctx, false, false, rows, n, np[r], alpha, buf[bsel], ldbuf, 0, bcol + p_off[r], ldb, 0, 1.0, oc, xy, 0, 1, 0));
            RB_CUDA(cudaEventRecord(ctx->aux_ev[3 + bsel], ctx->stream));
        }
    }
    return RB_OK;
}
