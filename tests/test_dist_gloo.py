"""world_size-2 gloo test of the multi-GPU host logic on CPU: the P-shard partition (shard_range ==
iter_auxbas(P_lo..P_hi), reference src/ri.rs:190-198) plus ONE all-reduce(sum) of the J and K partials reproduces the
unsharded result (likewise the (ia|jb) blocks built from the local rows of ri3mo); ao2mo and d_P need no
communication.  The per-rank partials come from the CPU oracle here (there is no GPU in this tier); the GPU tier runs the same flow with the CUDA kernels (tests/test_gpu_dist.py, bench.py --gpus N)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, nb, naux, no, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.api import Oracle
        from rest_tensors_b200.device import shard_range, all_reduce_sum, gather_dp
        o = Oracle()
        p_lo, p_hi = shard_range(naux, rank, world)
        nx = p_hi - p_lo
        ri_local = o.fill_ri3ao_symm(nb, p_lo, p_hi)           # this rank's slabs only
        c = o.fill_linear(nb * nb, 3, scale=nb ** -0.5)
        cm = c.reshape((nb, nb), order="F")
        dm = np.ascontiguousarray((2.0 * cm[:, :no] @ cm[:, :no].T).reshape(-1, order="F"))
        ct = np.ascontiguousarray((cm[:, :no] * np.sqrt(2.0)).reshape(-1, order="F"))
        d_local = o.ri_dp(ri_local, dm, nb, nx)
        j = torch.from_numpy(o.ri_j(ri_local, d_local, nb, nx))
        k = torch.from_numpy(o.ri_k(ri_local, ct, nb, no, nx))
        all_reduce_sum(j, world)
        all_reduce_sum(k, world)
        mo_local = o.ri_ao2mo_f(c, ri_local, nb, nb, nx)
        # (ia|jb) block from the local rows of ri3mo: a partial sum over P, completed by one all-reduce like J and K
        box_a, box_b = (0, no, no, nb - no), (1, no - 1, no, nb - no)
        g = torch.from_numpy(o.ri_iajb(nx, mo_local, nb, box_a, mo_local, nb, box_b))
        all_reduce_sum(g, world)
        # RPA-type block row out[P_local, all Q]: the host logic of ShardedRI.mo_pq (all-gather of the box's row blocks,
        # one GEMM per received block) with the per-block GEMM supplied by the oracle instead of the CUDA kernel
        from rest_tensors_b200.device import ShardedRI
        box = (0, no, no, nb - no)
        wts = o.fill_linear(box[1] * box[3], 7)

        class _OracleBlocks(ShardedRI):
            def _mo_pq_block(self, mine, piece, ldp, nq, ll, rl, w, out, q_lo):
                a = np.ascontiguousarray(mine.numpy().reshape((ldp, ll * rl), order="F")[: self.nx].reshape(-1, order="F"))
                b = np.ascontiguousarray(piece.numpy().reshape((ldp, ll * rl), order="F")[:nq].reshape(-1, order="F"))
                blk = o.ri_mo_pq(a, self.nx, b, nq, ll, (0, ll, 0, rl), w.numpy())
                out[q_lo * self.nx:(q_lo + nq) * self.nx] = torch.from_numpy(blk)

        shd = _OracleBlocks(None, nb, naux, rank, world, data=torch.from_numpy(ri_local))
        pq_rows = shd.mo_pq(torch.from_numpy(mo_local), nb, nb, box, torch.from_numpy(wts), out=torch.zeros(nx * naux, dtype=torch.float64))
        pq_all = [torch.zeros((naux - naux // world) * naux + naux, dtype=torch.float64) for _ in range(world)]
        pad = torch.zeros_like(pq_all[0]); pad[: pq_rows.numel()] = pq_rows
        dist.all_gather(pq_all, pad)
        # gather the d_P pieces and the P-rows of ri3mo for the check on rank 0
        d_full = gather_dp(torch.from_numpy(d_local), naux, p_lo, world)
        if rank == 0:
            ri = o.fill_ri3ao_symm(nb, 0, naux)
            d_ref = o.ri_dp(ri, dm, nb, naux)
            ok = np.allclose(d_full.numpy(), d_ref, rtol=1e-12, atol=0)
            ok &= np.allclose(j.numpy(), o.ri_j(ri, d_ref, nb, naux), rtol=1e-11, atol=1e-12)
            ok &= np.allclose(k.numpy(), o.ri_k(ri, ct, nb, no, naux), rtol=1e-11, atol=1e-12)
            mo = o.ri_ao2mo_f(c, ri, nb, nb, naux).reshape((naux, nb, nb), order="F")
            ok &= np.array_equal(mo[p_lo:p_hi].reshape(-1, order="F"), mo_local)
            mo_flat = np.ascontiguousarray(mo.reshape(-1, order="F"))
            ok &= np.allclose(g.numpy(), o.ri_iajb(naux, mo_flat, nb, box_a, mo_flat, nb, box_b), rtol=1e-11, atol=1e-12)
            pq_ref = o.ri_mo_pq(mo_flat, naux, mo_flat, naux, nb, box, wts).reshape((naux, naux), order="F")
            for s_ in range(world):
                lo, hi = shard_range(naux, s_, world)
                rows = pq_all[s_][: (hi - lo) * naux].numpy().reshape((hi - lo, naux), order="F")
                ok &= np.allclose(rows, pq_ref[lo:hi], rtol=1e-11, atol=1e-12)
            ret.put(bool(ok))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_p_sharding_with_allreduce_world2():
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 12, 9, 3, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert ret.get(timeout=5) is True
