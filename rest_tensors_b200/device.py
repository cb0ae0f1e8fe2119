"""Device-resident API: librest_b200's rb_* entry points on HBM buffers, plus the P-sharded RI tensor.

PyTorch is plumbing only here -- it owns device memory (``torch.empty(..., device='cuda')``), the current CUDA
stream and the NCCL process group.  Every kernel that runs is one of librest_b200.so's own sm_100a kernels,
called through the C ABI with raw device pointers.

Sharding: the auxiliary index P is the slowest index of ri3ao[nb, nb, naux], so rank r of G owns the
contiguous block P in [floor(r*naux/G), floor((r+1)*naux/G)) -- exactly the reference's ``iter_auxbas(range)``
(src/ri.rs:190-198).  ao2mo and d_P need no communication; J and K are partial sums over the local slabs and
are completed by ONE all-reduce(sum, f64) each over NVLink.  The collectives live behind the C ABI (rb_comm_* /
rb_allreduce_sum / rb_ri_j_allreduce / rb_ri_k_allreduce: NCCL bound by librest_b200 itself, on the context's stream), so
a Rust host gets the same multi-GPU path; torch.distributed only carries the 128-byte NCCL unique id and the host barriers
here.  CPU tensors (the gloo host-logic tests) go through torch.distributed.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from ._lib import lib, check, ch, RestB200Error  # noqa: F401


def shard_range(naux: int, rank: int, world: int) -> Tuple[int, int]:
    """P_lo = floor(r*naux/G), P_hi = floor((r+1)*naux/G)"""
    return (rank * naux) // world, ((rank + 1) * naux) // world


def _p(t) -> C.c_void_p:
    """device pointer of a tensor (None -> NULL); a raw integer address (e.g. an opened peer block) passes through"""
    if t is None:
        return C.c_void_p(0)
    if isinstance(t, int):
        return C.c_void_p(t)
    return C.c_void_p(t.data_ptr())


class Graph:
    """A recorded call sequence (rb_graph_end); launch() replays it on the context's stream."""

    def __init__(self, ctx: "Context", handle):
        self.ctx, self.h = ctx, handle

    @property
    def kernels(self) -> int:
        return int(lib.rb_graph_kernel_count(self.h))

    def launch(self) -> None:
        check(lib.rb_graph_launch(self.ctx.h, self.h), "rb_graph_launch")

    def close(self) -> None:
        if self.h and self.ctx.h:
            check(lib.rb_graph_free(self.ctx.h, self.h), "rb_graph_free")
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _Recording:
    def __init__(self, ctx: "Context"):
        self.ctx, self.graph = ctx, None

    def __enter__(self):
        check(lib.rb_graph_begin(self.ctx.h), "rb_graph_begin")
        return self

    def __exit__(self, exc_type, exc, tb):
        h = C.c_void_p()
        st = lib.rb_graph_end(self.ctx.h, C.byref(h))
        if exc_type is None:
            check(st, "rb_graph_end")
            self.graph = Graph(self.ctx, h)
        elif st == 0:
            lib.rb_graph_free(self.ctx.h, h)
        return False


class Context:
    """One rb_ctx bound to a CUDA device; calls run on torch's current stream for that device."""

    def __init__(self, device: Optional[int] = None):
        if not torch.cuda.is_available():
            raise RestB200Error("rest_tensors_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.cuda.current_device() if device is None else int(device)
        h = C.c_void_p()
        check(lib.rb_ctx_create(self.device, C.byref(h)), "rb_ctx_create")
        self.h = h
        self.bind_stream()

    def bind_stream(self) -> None:
        s = torch.cuda.current_stream(self.device)
        check(lib.rb_ctx_set_stream(self.h, C.c_void_p(s.cuda_stream)), "rb_ctx_set_stream")

    def close(self) -> None:
        if getattr(self, "h", None):
            lib.rb_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers --
    def empty(self, n: int) -> torch.Tensor:
        return torch.empty(int(n), dtype=torch.float64, device=f"cuda:{self.device}")

    def sync(self) -> None:
        check(lib.rb_ctx_sync(self.h), "rb_ctx_sync")

    @property
    def num_sms(self) -> int:
        return lib.rb_ctx_num_sms(self.h)

    @property
    def launches(self) -> int:
        return int(lib.rb_ctx_launch_count(self.h))

    @property
    def tma_layout_launches(self) -> int:
        """launches of the bulk-tensor (TMA) copy / transpose / pack / unpack kernels so far"""
        return int(lib.rb_ctx_tma_layout_count(self.h))

    def set_layout_path(self, path: int) -> None:
        """0: 32-byte / plain-load copy and transpose kernels (default), 1: bulk-tensor (TMA) kernels where eligible"""
        check(lib.rb_ctx_set_layout_path(self.h, path), "rb_ctx_set_layout_path")

    def set_gemm_path(self, path: int) -> None:
        check(lib.rb_ctx_set_gemm_path(self.h, path), "rb_ctx_set_gemm_path")

    def use_own_stream(self) -> None:
        """run this context's calls on the library's own (non-default) stream instead of torch's current one"""
        check(lib.rb_ctx_use_own_stream(self.h), "rb_ctx_use_own_stream")

    def record(self) -> "_Recording":
        """`with ctx.record() as rec: <calls>` records the calls on this context into a CUDA graph (they do not run); afterwards
        `rec.graph.launch()` replays them with one launch.  The context must be on a non-default stream (bind_stream() inside
        `torch.cuda.stream(s)`, or use_own_stream()), the sequence must have run once before, and every buffer it touches --
        outputs included -- must exist before recording (no torch allocation inside)."""
        return _Recording(self)

    def poison_workspaces(self):
        """Test hook: every internal workspace := NaN pattern (a stale read then shows in the result)."""
        check(lib.rb_ctx_poison_workspaces(self.h), "rb_ctx_poison_workspaces")

    # -- collectives (NCCL inside librest_b200) --
    @property
    def comm_world(self) -> int:
        return int(lib.rb_comm_world(self.h))

    @property
    def comm_rank(self) -> int:
        return int(lib.rb_comm_rank(self.h))

    def comm_init(self, rank: int, world: int) -> None:
        """Give this context a communicator over `world` ranks (one process per GPU).  Rank 0 draws the NCCL unique id
        inside the library; the 128 bytes reach the other ranks through the torch.distributed process group (any
        out-of-band channel would do: a Rust host uses a file or a pipe)."""
        if world <= 1 or self.comm_world == world:
            return
        import torch.distributed as dist
        ident = (C.c_ubyte * 128)()
        if rank == 0:
            check(lib.rb_comm_unique_id(ident), "rb_comm_unique_id")
        box = [bytes(ident)]
        dist.broadcast_object_list(box, src=0)
        ident = (C.c_ubyte * 128).from_buffer_copy(box[0])
        check(lib.rb_comm_init_rank(self.h, rank, world, ident), "rb_comm_init_rank")

    def comm_destroy(self) -> None:
        check(lib.rb_comm_destroy(self.h), "rb_comm_destroy")

    def allreduce_sum(self, t: torch.Tensor) -> None:
        check(lib.rb_allreduce_sum(self.h, _p(t), t.numel()), "rb_allreduce_sum")

    def allgather_shards(self, local: torch.Tensor, naux: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if out is None:
            out = self.empty(naux)
        check(lib.rb_allgather_shards(self.h, _p(local), _p(out), naux), "rb_allgather_shards")
        return out

    def nccl_version(self):
        v = C.c_int(0)
        path = C.create_string_buffer(256)
        check(lib.rb_comm_nccl_version(C.byref(v), path, 256), "rb_comm_nccl_version")
        return int(v.value), path.value.decode()

    # -- BLAS --
    def dgemm(self, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc) -> None:
        check(lib.rb_dgemm(self.h, ch(ta), ch(tb), m, n, k, alpha, _p(a), lda, _p(b), ldb, beta, _p(c), ldc), "rb_dgemm")

    def dgemm_strided_batched(self, ta, tb, m, n, k, alpha, a, lda, sa, b, ldb, sb, beta, c, ldc, sc, batch) -> None:
        check(lib.rb_dgemm_strided_batched(self.h, ch(ta), ch(tb), m, n, k, alpha, _p(a), lda, sa, _p(b), ldb, sb, beta,
                                           _p(c), ldc, sc, batch), "rb_dgemm_strided_batched")

    def dsyrk(self, uplo, trans, n, k, alpha, a, lda, beta, c, ldc) -> None:
        check(lib.rb_dsyrk(self.h, ch(uplo), ch(trans), n, k, alpha, _p(a), lda, beta, _p(c), ldc), "rb_dsyrk")

    def dgemv(self, trans, m, n, alpha, a, lda, x, incx, beta, y, incy) -> None:
        check(lib.rb_dgemv(self.h, ch(trans), m, n, alpha, _p(a), lda, _p(x), incx, beta, _p(y), incy), "rb_dgemv")

    def dsymm(self, side, uplo, m, n, alpha, a, lda, b, ldb, beta, c, ldc) -> None:
        check(lib.rb_dsymm(self.h, ch(side), ch(uplo), m, n, alpha, _p(a), lda, _p(b), ldb, beta, _p(c), ldc), "rb_dsymm")

    # -- RI contractions --
    def ri_ao2mo(self, c_left, nl, c_right, nr, ri3ao, out, nb, nx, out_ldp=None) -> None:
        check(lib.rb_ri_ao2mo(self.h, _p(c_left), nl, _p(c_right), nr, _p(ri3ao), _p(out), nb, nx,
                              nx if out_ldp is None else out_ldp), "rb_ri_ao2mo")

    def ri_dp(self, ri3ao, dm, d, nb, nx) -> None:
        check(lib.rb_ri_dp(self.h, _p(ri3ao), _p(dm), _p(d), nb, nx), "rb_ri_dp")

    def ri_j(self, ri3ao, d, j, nb, nx) -> None:
        check(lib.rb_ri_j(self.h, _p(ri3ao), _p(d), _p(j), nb, nx), "rb_ri_j")

    def ri_dp_j(self, ri3ao, dm, d, j, nb, nx) -> None:
        """d_P and J from one read of ri3ao (rb_ri_dp_j: persistent cooperative kernel; falls back to the two GEMVs where it cannot run)"""
        check(lib.rb_ri_dp_j(self.h, _p(ri3ao), _p(dm), _p(d), _p(j), nb, nx), "rb_ri_dp_j")

    def ri_k(self, ri3ao, ct, no, k, nb, nx) -> None:
        check(lib.rb_ri_k(self.h, _p(ri3ao), _p(ct), no, _p(k), nb, nx), "rb_ri_k")

    def ri_iajb(self, np_, mo_a, ldp_a, nl_a, nr_a, box_a, mo_b, ldp_b, nl_b, nr_b, box_b, beta, out, ldo) -> None:
        """box = (l0, ll, r0, rl)"""
        check(lib.rb_ri_iajb(self.h, np_, _p(mo_a), ldp_a, nl_a, nr_a, *box_a, _p(mo_b), ldp_b, nl_b, nr_b, *box_b, beta,
                             _p(out), ldo), "rb_ri_iajb")

    def ri_mo_pq(self, mo_a, ldp_a, np_a, mo_b, ldp_b, np_b, nl, nr, box, w, beta, out, ldo) -> None:
        check(lib.rb_ri_mo_pq(self.h, _p(mo_a), ldp_a, np_a, _p(mo_b), ldp_b, np_b, nl, nr, *box, _p(w), beta, _p(out),
                              ldo), "rb_ri_mo_pq")

    def special_dgemm_01(self, ten3, x_a, y_a, z_a, sx, lx, sz, lz, b, ldb, lcb, alpha, beta) -> None:
        check(lib.rb_special_dgemm_01(self.h, _p(ten3), x_a, y_a, z_a, sx, lx, sz, lz, _p(b), ldb, lcb, alpha, beta),
              "rb_special_dgemm_01")

    # -- eigen-solvers (one-sided Jacobi; these synchronise) --
    def dsyev(self, jobz, uplo, n, a, lda, w, z, ldz) -> None:
        check(lib.rb_dsyev(self.h, ch(jobz), ch(uplo), n, _p(a), lda, _p(w), _p(z), ldz), "rb_dsyev")

    def dspev(self, n, ap, w, z, ldz) -> None:
        check(lib.rb_dspev(self.h, n, _p(ap), _p(w), _p(z), ldz), "rb_dspev")

    def dspgv(self, n, ap, bp, m, w, z, ldz) -> None:
        check(lib.rb_dspgv(self.h, n, _p(ap), _p(bp), m, _p(w), _p(z), ldz), "rb_dspgv")

    def matrix_power(self, n, a, lda, p, threshold, out, ldo) -> int:
        kept = C.c_int(0)
        check(lib.rb_matrix_power(self.h, n, _p(a), lda, p, threshold, _p(out), ldo, C.byref(kept)), "rb_matrix_power")
        return kept.value

    # -- layout --
    def pack_upper(self, full, n, packed) -> None:
        check(lib.rb_pack_upper(self.h, _p(full), n, _p(packed)), "rb_pack_upper")

    def unpack_upper(self, packed, n, full) -> None:
        check(lib.rb_unpack_upper(self.h, _p(packed), n, _p(full)), "rb_unpack_upper")

    def ri_pack_symm(self, ri, nao, naux, out) -> None:
        check(lib.rb_ri_pack_symm(self.h, _p(ri), nao, naux, _p(out)), "rb_ri_pack_symm")

    def copy_mm(self, xl, yl, f, fx, fy, fxs, fys, t, tx, ty, txs, tys) -> None:
        check(lib.rb_copy_mm(self.h, xl, yl, _p(f), fx, fy, fxs, fys, _p(t), tx, ty, txs, tys), "rb_copy_mm")

    def copy_mr(self, xl, yl, f, fx, fy, fxs, fys, t, tx, ty, tz, txs, tys, t3, mod) -> None:
        check(lib.rb_copy_mr(self.h, xl, yl, _p(f), fx, fy, fxs, fys, _p(t), tx, ty, tz, txs, tys, t3, mod), "rb_copy_mr")

    def copy_rm(self, xl, yl, f, fx, fy, fz, fxs, fys, f3, mod, t, tx, ty, txs, tys) -> None:
        check(lib.rb_copy_rm(self.h, xl, yl, _p(f), fx, fy, fz, fxs, fys, f3, mod, _p(t), tx, ty, txs, tys), "rb_copy_rm")

    def copy_rr(self, xl, yl, zl, f, fx, fy, fz, fxs, fys, fzs, t, tx, ty, tz, txs, tys, tzs) -> None:
        check(lib.rb_copy_rr(self.h, xl, yl, zl, _p(f), fx, fy, fz, fxs, fys, fzs, _p(t), tx, ty, tz, txs, tys, tzs),
              "rb_copy_rr")

    def erifold4_chunk_copy(self, eri, size0, size1, ld, ranges, buf, mode) -> None:
        """ERIFold4 chunk scatter on device buffers (src/eri.rs:266-372); ranges = ((i0, i1), (j0, j1), (k0, k1), (l0, l1))"""
        (i0, i1), (j0, j1), (k0, k1), (l0, l1) = ranges
        check(lib.rb_erifold4_chunk_copy(self.h, _p(eri), size0, size1, ld, i0, i1 - i0, j0, j1 - j0, k0, k1 - k0, l0, l1 - l0,
                                         _p(buf), mode), "rb_erifold4_chunk_copy")

    def ri_transpose(self, inp, i, j, k, which, out) -> None:
        check(lib.rb_ri_transpose(self.h, _p(inp), i, j, k, which, _p(out)), "rb_ri_transpose")

    def matrix_transpose(self, inp, rows, cols, out) -> None:
        check(lib.rb_matrix_transpose(self.h, _p(inp), rows, cols, _p(out)), "rb_matrix_transpose")

    def self_scaled_add(self, c, p, b, n) -> None:
        check(lib.rb_self_scaled_add(self.h, _p(c), _p(p), b, n), "rb_self_scaled_add")

    def self_general_add(self, c, p, a, b, n) -> None:
        check(lib.rb_self_general_add(self.h, _p(c), _p(p), a, b, n), "rb_self_general_add")

    def self_multiple(self, c, a, n) -> None:
        check(lib.rb_self_multiple(self.h, _p(c), a, n), "rb_self_multiple")

    def self_add(self, c, p, n) -> None:
        check(lib.rb_self_add(self.h, _p(c), _p(p), n), "rb_self_add")

    def self_sub(self, c, p, n) -> None:
        check(lib.rb_self_sub(self.h, _p(c), _p(p), n), "rb_self_sub")

    # -- synthetic inputs / probes --
    def fill_linear(self, v, n, seed, idx0, scale) -> None:
        check(lib.rb_fill_linear(self.h, _p(v), n, seed, idx0, scale), "rb_fill_linear")

    def fill_ri3ao_symm(self, a, nb, p_lo, p_hi, seed, scale) -> None:
        check(lib.rb_fill_ri3ao_symm(self.h, _p(a), nb, p_lo, p_hi, seed, scale), "rb_fill_ri3ao_symm")

    def fp64_peak_probe(self, kind: int, iters: int = 4096):
        tf, ms = C.c_double(), C.c_double()
        check(lib.rb_fp64_peak_probe(self.h, kind, iters, C.byref(tf), C.byref(ms)), "rb_fp64_peak_probe")
        return tf.value, ms.value

    def hbm_copy_probe(self, nbytes: int, iters: int = 10) -> float:
        g = C.c_double()
        check(lib.rb_hbm_copy_probe(self.h, nbytes, iters, C.byref(g)), "rb_hbm_copy_probe")
        return g.value


class ShardedRI:
    """One rank's P-shard of ri3ao[nb, nb, naux], resident in HBM across calls (the SCF loop re-uses it)."""

    def __init__(self, ctx: Context, nb: int, naux: int, rank: int = 0, world: int = 1,
                 data: Optional[torch.Tensor] = None, comm: bool = True):
        """comm=False: do not set up a communicator here (a single process that drives several contexts calls
        rb_comm_init_all itself and brackets the collectives with rb_comm_group_start / _end)"""
        self.ctx, self.nb, self.naux, self.rank, self.world = ctx, int(nb), int(naux), int(rank), int(world)
        self.p_lo, self.p_hi = shard_range(self.naux, self.rank, self.world)
        self.nx = self.p_hi - self.p_lo
        n = self.nb * self.nb * self.nx
        if data is None:
            data = ctx.empty(n)
        elif data.numel() < n:
            raise ValueError("ShardedRI: buffer smaller than the local shard")
        self.data = data
        if comm and ctx is not None and self.world > 1 and data.is_cuda:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():   # one process per GPU; otherwise the host wires the comm itself
                ctx.comm_init(self.rank, self.world)

    def fill_synthetic(self, seed: int = 1, scale: float = 1.0) -> "ShardedRI":
        self.ctx.fill_ri3ao_symm(self.data, self.nb, self.p_lo, self.p_hi, seed, scale)
        return self

    def ao2mo(self, c_left: torch.Tensor, nl: int, c_right: torch.Tensor, nr: int,
              out: Optional[torch.Tensor] = None, out_ldp: Optional[int] = None) -> torch.Tensor:
        """local rows ri3mo[P_lo..P_hi, :, :] as a dense [nx_local, nl, nr] column-major buffer (no communication).
        With out_ldp > nx_local the rows land in a wider P-fastest tensor (out already offset to this shard's first P):
        that is how a rank's pitched sub-array of the global ri3mo is written in place."""
        if out is None:
            out = self.ctx.empty(self.nx * nl * nr)
        self.ctx.ri_ao2mo(c_left, nl, c_right, nr, self.data, out, self.nb, self.nx, out_ldp)
        return out

    def dp(self, dm: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """d[P_lo..P_hi] from the replicated density (no communication needed before J)"""
        if out is None:
            out = self.ctx.empty(self.nx)
        self.ctx.ri_dp(self.data, dm, out, self.nb, self.nx)
        return out

    def j(self, d_local: torch.Tensor, out: Optional[torch.Tensor] = None, reduce: bool = True) -> torch.Tensor:
        if out is None:
            out = self.ctx.empty(self.nb * self.nb)
        if reduce and self.world > 1:   # partial sum + NCCL all-reduce on the same stream, one C-ABI call
            check(lib.rb_ri_j_allreduce(self.ctx.h, _p(self.data), _p(d_local), _p(out), self.nb, self.nx), "rb_ri_j_allreduce")
        else:
            self.ctx.ri_j(self.data, d_local, out, self.nb, self.nx)
        return out

    def dp_j(self, dm: torch.Tensor, out_d: Optional[torch.Tensor] = None, out_j: Optional[torch.Tensor] = None,
             reduce: bool = True):
        """(d[P_lo..P_hi), J) from ONE pass over the shard; J is all-reduced over the ranks when reduce (rb_allreduce_sum)"""
        if out_d is None:
            out_d = self.ctx.empty(self.nx)
        if out_j is None:
            out_j = self.ctx.empty(self.nb * self.nb)
        self.ctx.ri_dp_j(self.data, dm, out_d, out_j, self.nb, self.nx)
        if reduce and self.world > 1:
            self.ctx.allreduce_sum(out_j)
        return out_d, out_j

    def k(self, ct: torch.Tensor, no: int, out: Optional[torch.Tensor] = None, reduce: bool = True) -> torch.Tensor:
        if out is None:
            out = self.ctx.empty(self.nb * self.nb)
        if reduce and self.world > 1:
            check(lib.rb_ri_k_allreduce(self.ctx.h, _p(self.data), _p(ct), no, _p(out), self.nb, self.nx), "rb_ri_k_allreduce")
        else:
            self.ctx.ri_k(self.data, ct, no, out, self.nb, self.nx)
        return out

    def gather_dp(self, d_local: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """full d[0..naux) on every rank (rb_allgather_shards)"""
        return self.ctx.allgather_shards(d_local, self.naux, out)

    def iajb(self, mo: torch.Tensor, nl: int, nr: int, box_a, box_b, mo_b: Optional[torch.Tensor] = None,
             nl_b: Optional[int] = None, nr_b: Optional[int] = None, out: Optional[torch.Tensor] = None,
             reduce: bool = True) -> torch.Tensor:
        """(ia|jb)-type block from this rank's rows of ri3mo (the dense [nx_local, nl, nr] buffer ao2mo returned): a
        partial sum over the local P, completed by ONE all-reduce like J and K.  box = (l0, ll, r0, rl)."""
        if mo_b is None:
            mo_b, nl_b, nr_b = mo, nl, nr
        m, n = box_a[1] * box_a[3], box_b[1] * box_b[3]
        if out is None:
            out = self.ctx.empty(m * n)
        self.ctx.ri_iajb(self.nx, mo, self.nx, nl, nr, box_a, mo_b, self.nx, nl_b, nr_b, box_b, 0.0, out, m)
        if reduce:
            all_reduce_sum(out, self.world, self.ctx)
        return out

    def mo_pq(self, mo: torch.Tensor, nl: int, nr: int, box, w: Optional[torch.Tensor] = None,
              out: Optional[torch.Tensor] = None, exchange: str = "auto") -> torch.Tensor:
        """RPA-type block row out[P_local, Q] = sum_{(l,r) in box} w[l,r] mo[P,l,r] mo[Q,l,r] over ALL Q in [0, naux):
        the one exchange step on this side of the path -- every rank needs every other rank's rows of ri3mo for the
        box.  out is [nx_local, naux], column-major.  world == 1: one symmetric GEMM, no communication.

        exchange = "p2p" (default on GPUs): each rank gathers its box into a panel it exposes to its peers (CUDA IPC
        over NVLink / NVSwitch) and rb_ri_mo_pq_peers runs the all-gather and the GEMMs as one pipeline: the copy
        engines pull peer s+1's panel on a second stream while the DMMA GEMM on peer s's panel runs.
        exchange = "allgather": NCCL all-gather of the panels, then one GEMM per received block (also the gloo / CPU
        host-logic tier)."""
        if out is None:
            out = self.ctx.empty(self.nx * self.naux)
        if self.world == 1:
            self.ctx.ri_mo_pq(mo, self.nx, self.nx, mo, self.nx, self.nx, nl, nr, box, w, 0.0, out, self.nx)
            return out
        import torch.distributed as dist
        l0, ll, r0, rl = box
        nx_max = -(-self.naux // self.world)
        nx_max += nx_max & 1                        # even pitch: TMA-describable panels
        if exchange == "auto":
            exchange = "p2p" if mo.is_cuda else "allgather"
        if exchange == "p2p":
            panels = self._peer_panels(nx_max * ll * rl * 8)
            check(lib.rb_copy_rr(self.ctx.h, self.nx, ll, rl, _p(mo), self.nx, nl, nr, 0, l0, r0, _p(panels.local), nx_max,
                                 ll, rl, 0, 0, 0), "rb_copy_rr")
            self.ctx.sync()
            dist.barrier()                          # every rank's panel is complete and visible
            ranges = [shard_range(self.naux, s, self.world) for s in range(self.world)]
            ptrs = (C.c_void_p * self.world)(*panels.ptrs)
            rows = (C.c_int * self.world)(*[hi - lo for lo, hi in ranges])
            offs = (C.c_int64 * self.world)(*[lo for lo, _ in ranges])
            check(lib.rb_ri_mo_pq_peers(self.ctx.h, self.rank, self.world, ptrs, nx_max, rows, ll * rl, _p(w), _p(out),
                                        self.nx, offs), "rb_ri_mo_pq_peers")
            self.ctx.sync()
            dist.barrier()                          # nobody rewrites its panel while a peer still reads it
            return out
        mine = torch.zeros(nx_max * ll * rl, dtype=mo.dtype, device=mo.device)
        if mo.is_cuda:  # gather this rank's box columns into a dense [nx_max, ll*rl] panel with our own copy kernel
            check(lib.rb_copy_rr(self.ctx.h, self.nx, ll, rl, _p(mo), self.nx, nl, nr, 0, l0, r0, _p(mine), nx_max, ll, rl,
                                 0, 0, 0), "rb_copy_rr")
        else:           # gloo tier (host logic test): same gather with tensor views
            mine.view(rl, ll, nx_max)[:, :, : self.nx] = mo.view(nr, nl, self.nx)[r0:r0 + rl, l0:l0 + ll, :]
        pieces = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(pieces, mine)
        for s, piece in enumerate(pieces):
            q_lo, q_hi = shard_range(self.naux, s, self.world)
            self._mo_pq_block(mine, piece, nx_max, q_hi - q_lo, ll, rl, w, out, q_lo)
        return out

    def special_dgemm_p(self, b: torch.Tensor, alpha: float = 1.0, beta: float = 0.0) -> "ShardedRI":
        """special_dgemm_f_01 over the SHARDED index (restmatr.f90:111-154 with the full x range, every y, z = P):
        T[:, :, P'] <- alpha * sum_P T[:, :, P] b[P, P'] + beta * T[:, :, P'], b = [naux, naux] replicated (e.g. V^-1/2
        when ri3ao is built from the raw three-centre integrals).  In place in the reference's sense:
        afterwards self.data holds the transformed shard.  world > 1: every rank needs every other rank's shard; they are
        pulled chunk by chunk over NVLink by rb_special_dgemm_01_peers while the GEMMs run (no NCCL on the data path)."""
        xy = self.nb * self.nb
        if self.world == 1:
            self.ctx.special_dgemm_01(self.data, self.nb, self.nb, self.nx, 0, self.nb, 0, self.nx, b, self.naux, self.nx,
                                      alpha, beta)
            return self
        import torch.distributed as dist
        views = PeerViews(self.ctx, self.data, self.rank, self.world)
        out = self.ctx.empty(xy * self.nx)
        ranges = [shard_range(self.naux, s, self.world) for s in range(self.world)]
        ptrs = (C.c_void_p * self.world)(*views.ptrs)
        cols = (C.c_int * self.world)(*[hi - lo for lo, hi in ranges])
        offs = (C.c_int64 * self.world)(*[lo for lo, _ in ranges])
        self.ctx.sync()
        dist.barrier()                              # every shard is complete and visible
        check(lib.rb_special_dgemm_01_peers(self.ctx.h, self.rank, self.world, ptrs, xy, cols, offs, _p(b), self.naux, alpha,
                                            beta, _p(out)), "rb_special_dgemm_01_peers")
        views.close()                               # sync + barrier: nobody is still reading the old shard
        self.data = out
        return self

    def _peer_panels(self, nbytes: int) -> "PeerBlocks":
        cur = getattr(self, "_panels", None)
        if cur is None or cur.nbytes < nbytes:      # same size on every rank, so all ranks re-create together
            if cur is not None:
                cur.close()
            self._panels = PeerBlocks(self.ctx, nbytes, self.rank, self.world)
        return self._panels

    def _mo_pq_block(self, mine, piece, ldp, nq, ll, rl, w, out, q_lo) -> None:
        self.ctx.ri_mo_pq(mine, ldp, self.nx, piece, ldp, nq, ll, rl, (0, ll, 0, rl), w, 0.0,
                          out[q_lo * self.nx:], self.nx)


_IPC_OPEN = {}  # handle bytes -> [mapped base address, reference count]: a handle is opened once per process


def _ipc_open(ctx: "Context", handle: bytes) -> int:
    ent = _IPC_OPEN.get(handle)
    if ent is None:
        p = C.c_void_p()
        check(lib.rb_ipc_open(ctx.h, (C.c_ubyte * 64).from_buffer_copy(handle), C.byref(p)), "rb_ipc_open")
        ent = _IPC_OPEN[handle] = [int(p.value), 0]
    ent[1] += 1
    return ent[0]


def _ipc_close(ctx: "Context", handle: bytes) -> None:
    ent = _IPC_OPEN.get(handle)
    if ent is None:
        return
    ent[1] -= 1
    if ent[1] <= 0:
        check(lib.rb_ipc_close(ctx.h, C.c_void_p(ent[0])), "rb_ipc_close")
        del _IPC_OPEN[handle]


class PeerViews:
    """Every rank's device buffer `local` (any allocation: a torch tensor's storage or an rb_dev_alloc block) mapped into
    every rank's address space.  The owner exports the allocation that holds the buffer as a 64-byte CUDA IPC handle plus
    the buffer's offset inside it, the (handle, offset) pairs travel through the process group, and every rank opens its
    peers' handles.  ptrs[s] is the address of rank s's buffer as seen from this rank (own buffer: the local pointer);
    any rb_* device entry point and cudaMemcpy can read it over NVLink."""

    def __init__(self, ctx: Context, local, rank: int, world: int):
        import torch.distributed as dist
        self.ctx, self.rank, self.world = ctx, rank, world
        self.local = local if isinstance(local, int) else int(local.data_ptr())
        handle = (C.c_ubyte * 64)()
        off = C.c_int64(0)
        check(lib.rb_ipc_export(ctx.h, C.c_void_p(self.local), handle, C.byref(off)), "rb_ipc_export")
        pairs = [None] * world
        dist.all_gather_object(pairs, (bytes(handle), int(off.value)))
        self.handles = [h for h, _ in pairs]
        self.ptrs = []
        for s, (hb, o) in enumerate(pairs):
            self.ptrs.append(self.local if s == rank else _ipc_open(ctx, hb) + o)

    def close(self) -> None:
        import torch.distributed as dist
        self.ctx.sync()
        dist.barrier()                              # no peer is still reading
        for s, hb in enumerate(self.handles):
            if s != self.rank:
                _ipc_close(self.ctx, hb)
        dist.barrier()                              # every mapping is gone before an owner may free
        self.ptrs, self.handles = [], []


class PeerBlocks(PeerViews):
    """One block of `nbytes` per rank, allocated with rb_dev_alloc (plain cudaMalloc) and mapped on every rank."""

    def __init__(self, ctx: Context, nbytes: int, rank: int, world: int):
        self.nbytes = int(nbytes)
        base = C.c_void_p()
        check(lib.rb_dev_alloc(ctx.h, self.nbytes, C.byref(base)), "rb_dev_alloc")
        super().__init__(ctx, int(base.value), rank, world)

    def close(self) -> None:
        local = self.local
        super().close()
        check(lib.rb_dev_free(self.ctx.h, C.c_void_p(local)), "rb_dev_free")
        self.local = 0


def all_reduce_sum(t: torch.Tensor, world: int, ctx: Optional[Context] = None) -> None:
    """The only collective on the path: sum of the per-rank J / K partials.  Device tensors with a context that holds a
    communicator: rb_allreduce_sum (NCCL inside librest_b200, on the context's stream).  CPU tensors (gloo host-logic
    tier) or no context: torch.distributed."""
    if world <= 1:
        return
    if ctx is not None and t.is_cuda and ctx.comm_world == world:
        ctx.allreduce_sum(t)
        return
    import torch.distributed as dist
    dist.all_reduce(t, op=dist.ReduceOp.SUM)


def gather_dp(d_local: torch.Tensor, naux: int, p_lo: int, world: int, ctx: Optional[Context] = None) -> torch.Tensor:
    """Full d[0..naux) on every rank (38 KB at naux=4800).  With a communicator: rb_allgather_shards.  Otherwise each rank
    drops its piece into a zero vector and the pieces are summed -- shards may differ in length by one, which all_gather
    does not accept on every backend."""
    if ctx is not None and d_local.is_cuda and ctx.comm_world == world and world > 1:
        return ctx.allgather_shards(d_local, naux)
    full = torch.zeros(int(naux), dtype=d_local.dtype, device=d_local.device)
    full[p_lo:p_lo + d_local.numel()] = d_local
    all_reduce_sum(full, world)
    return full
