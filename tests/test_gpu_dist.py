"""Multi-GPU parity (needs >= 2 GPUs; skipped otherwise): each rank holds its P-shard of a synthetic ri3ao in HBM, runs
the CUDA kernels on it, and ONE all-reduce(sum) each completes J and K -- through the C ABI's own collectives
(rb_comm_init_rank / rb_ri_j_allreduce / rb_ri_k_allreduce / rb_allgather_shards: NCCL bound inside librest_b200).
Checked against the unsharded CPU oracle."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, nb, naux, no, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    try:
        from oracle.api import Oracle
        from rest_tensors_b200.device import Context, ShardedRI, gather_dp
        o = Oracle(); o.load_openblas()
        ctx = Context(rank)
        sh = ShardedRI(ctx, nb, naux, rank, world).fill_synthetic()
        c = o.fill_linear(nb * nb, 3, scale=nb ** -0.5)
        cm = c.reshape((nb, nb), order="F")
        dm = np.ascontiguousarray((2.0 * cm[:, :no] @ cm[:, :no].T).reshape(-1, order="F"))
        ct = np.ascontiguousarray((cm[:, :no] * np.sqrt(2.0)).reshape(-1, order="F"))
        dev = lambda a: torch.from_numpy(a).to(f"cuda:{rank}")  # noqa: E731
        cd = dev(c)
        d_local = sh.dp(dev(dm))
        j = sh.j(d_local)                  # partial + all-reduce
        k = sh.k(dev(ct), no)              # partial + all-reduce
        mo_local = sh.ao2mo(cd, nb, cd, nb)
        assert ctx.comm_world == world and ctx.comm_rank == rank      # the communicator lives inside librest_b200
        d_full = gather_dp(d_local, naux, sh.p_lo, world, ctx)        # rb_allgather_shards
        assert torch.equal(d_full, sh.gather_dp(d_local))
        j2 = sh.j(d_local, reduce=False); ctx.allreduce_sum(j2)       # partial + explicit rb_allreduce_sum == fused call
        assert torch.equal(j, j2)
        # consumers of ri3mo: (ia|jb) block (partial + all-reduce) and the RPA-type block row (all-gather of row blocks)
        box_a, box_b = (0, no, no, nb - no), (1, no - 1, no, nb - no)
        g = sh.iajb(mo_local, nb, nb, box_a, box_b)
        wts = o.fill_linear(box_a[1] * box_a[3], 7)
        pq = sh.mo_pq(mo_local, nb, nb, box_a, dev(wts))                              # peer panels over NVLink (default)
        pq_ag = sh.mo_pq(mo_local, nb, nb, box_a, dev(wts), exchange="allgather")      # NCCL all-gather of the panels
        pq_nw = sh.mo_pq(mo_local, nb, nb, box_a, None)                               # no weights: TMA reads the peer panel
        pq_nw_ag = sh.mo_pq(mo_local, nb, nb, box_a, None, exchange="allgather")
        torch.cuda.synchronize()
        ri = o.fill_ri3ao_symm(nb, 0, naux)
        d_ref = o.ri_dp(ri, dm, nb, naux)

        def err(x, y):
            return float(np.max(np.abs(x - y)) / np.max(np.abs(y)))
        mo_all = o.ri_ao2mo_f(c, ri, nb, nb, naux)
        mo_ref = mo_all.reshape((naux, nb, nb), order="F")[sh.p_lo:sh.p_hi].reshape(-1, order="F")
        g_ref = o.ri_iajb(naux, mo_all, nb, box_a, mo_all, nb, box_b)
        pq_ref = o.ri_mo_pq(mo_all, naux, mo_all, naux, nb, box_a, wts).reshape((naux, naux), order="F")[sh.p_lo:sh.p_hi]
        pq_nw_ref = o.ri_mo_pq(mo_all, naux, mo_all, naux, nb, box_a, None).reshape((naux, naux), order="F")[sh.p_lo:sh.p_hi]
        errs = [err(d_full.cpu().numpy(), d_ref), err(j.cpu().numpy(), o.ri_j(ri, d_ref, nb, naux)),
                err(k.cpu().numpy(), o.ri_k(ri, ct, nb, no, naux)), err(mo_local.cpu().numpy(), mo_ref),
                err(g.cpu().numpy(), g_ref), err(pq.cpu().numpy(), pq_ref.reshape(-1, order="F")),
                err(pq_ag.cpu().numpy(), pq_ref.reshape(-1, order="F")),
                err(pq_nw.cpu().numpy(), pq_nw_ref.reshape(-1, order="F")),
                err(pq_nw_ag.cpu().numpy(), pq_nw_ref.reshape(-1, order="F"))]
        # the step BEFORE the hot path: T <- alpha T B + beta T contracted over the SHARDED index,
        # peers' shards pulled over NVLink while the GEMMs run; afterwards sh.data is the transformed shard
        bmat = o.fill_linear(naux * naux, 8, scale=naux ** -0.5)
        t_ref = ri.copy()
        o.special_dgemm_f_01(t_ref, [nb, nb, naux], (0, nb), 0, (0, naux), bmat, [naux, naux], (0, naux), (0, naux), 0.7, -0.2)
        sh.special_dgemm_p(dev(bmat), 0.7, -0.2)
        torch.cuda.synchronize()
        errs.append(err(sh.data.cpu().numpy(), t_ref[sh.p_lo * nb * nb: sh.p_hi * nb * nb]))
        sh.special_dgemm_p(dev(bmat), 1.0, 0.0)       # a second pass on the transformed shards (beta = 0 branch)
        o.special_dgemm_f_01(t_ref, [nb, nb, naux], (0, nb), 0, (0, naux), bmat, [naux, naux], (0, naux), (0, naux), 1.0, 0.0)
        torch.cuda.synchronize()
        errs.append(err(sh.data.cpu().numpy(), t_ref[sh.p_lo * nb * nb: sh.p_hi * nb * nb]))
        # d_P and J from one read of the shard (rb_ri_dp_j), J completed by the same all-reduce
        os.environ["REST_B200_DPJ_FUSED"] = "1"
        sh3 = ShardedRI(ctx, nb, naux, rank, world).fill_synthetic()
        d1p, j1p = sh3.dp_j(dev(dm))
        del os.environ["REST_B200_DPJ_FUSED"]
        torch.cuda.synchronize()
        errs.append(err(j1p.cpu().numpy(), o.ri_j(ri, d_ref, nb, naux)))
        errs.append(err(gather_dp(d1p, naux, sh3.p_lo, world, ctx).cpu().numpy(), d_ref))
        # one launch per SCF iteration at N > 1: d_P + J + K with both all-reduces recorded into a CUDA graph on every rank
        side = torch.cuda.Stream()
        torch.cuda.synchronize()
        with torch.cuda.stream(side):
            ctx.bind_stream()
            dmd, ctd = dev(dm), dev(ct)
            sh2 = ShardedRI(ctx, nb, naux, rank, world).fill_synthetic()      # sh.data was transformed above
            dg, jg, kg = ctx.empty(sh2.nx), ctx.empty(nb * nb), ctx.empty(nb * nb)

            def iteration():
                sh2.dp(dmd, out=dg); sh2.j(dg, out=jg); sh2.k(ctd, no, out=kg)
            iteration(); side.synchronize()
            want_j, want_k = jg.clone(), kg.clone()
            assert torch.equal(want_j, j) and torch.equal(want_k, k)
            with ctx.record() as rec:
                iteration()
            for _ in range(3):
                jg.zero_(); kg.zero_()
                rec.graph.launch(); side.synchronize()
                assert torch.equal(jg, want_j) and torch.equal(kg, want_k), "graph replay with all-reduces differs"
            rec.graph.close()
        ctx.bind_stream()
        ret.put((rank, max(errs)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_jk_allreduce(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29600 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, 64, 203, 9, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    res = [ret.get(timeout=5) for _ in range(world)]
    assert all(e <= 1e-10 for _, e in res), res


def test_one_process_two_devices():
    """Contexts are independent: one host process drives two GPUs (a Rust host with a thread per device does this),
    launches overlap freely, and each device keeps its own tile scheduler / workspaces / per-device kernel attributes."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, ROOT)
    from oracle.api import Oracle
    from rest_tensors_b200.device import Context, ShardedRI
    o = Oracle(); o.load_openblas()
    nb, naux, no = 72, 160, 9
    c = o.fill_linear(nb * nb, 3, scale=nb ** -0.5)
    ri = o.fill_ri3ao_symm(nb, 0, naux)
    mo_ref = o.ri_ao2mo_f(c, ri, nb, nb, naux).reshape((naux, nb, nb), order="F")
    ctxs = [Context(0), Context(1)]
    shards = [ShardedRI(ctxs[r], nb, naux, r, 2).fill_synthetic() for r in range(2)]
    outs = []
    for rep in range(3):                      # interleave launches on the two devices without synchronising
        outs = []
        for r in range(2):
            cd = torch.from_numpy(c).to(f"cuda:{r}")
            outs.append(shards[r].ao2mo(cd, nb, cd, nb))
    for r in range(2):
        torch.cuda.synchronize(r)
        ref = mo_ref[shards[r].p_lo:shards[r].p_hi].reshape(-1, order="F")
        got = outs[r].cpu().numpy()
        assert float(np.max(np.abs(got - ref)) / np.max(np.abs(ref))) <= 1e-10, f"device {r}"


def test_one_process_peer_operand():
    """rb_peer_enable: a kernel of device 0's context reads an operand that lives in device 1's HBM (TMA loads over
    NVLink) and gives bitwise the result of the same call on a local copy."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, ROOT)
    from rest_tensors_b200.device import Context
    from rest_tensors_b200._lib import lib, check
    c0, c1 = Context(0), Context(1)
    check(lib.rb_peer_enable(c0.h, 1), "rb_peer_enable")
    m, n, k = 200, 136, 2000
    with torch.cuda.device(1):
        b_remote = c1.empty(n * k); c1.fill_linear(b_remote, n * k, 5, 0, 1.0)
        torch.cuda.synchronize(1)
    with torch.cuda.device(0):
        a = c0.empty(m * k); c0.fill_linear(a, m * k, 4, 0, 1.0)
        b_local = c0.empty(n * k); c0.fill_linear(b_local, n * k, 5, 0, 1.0)
        o1, o2 = c0.empty(m * n), c0.empty(m * n)
        c0.dgemm("N", "T", m, n, k, 1.0, a, m, b_local, n, 0.0, o1, m)
        c0.dgemm("N", "T", m, n, k, 1.0, a, m, b_remote, n, 0.0, o2, m)
        torch.cuda.synchronize(0)
        assert torch.equal(o1, o2)


def test_one_process_comm_init_all():
    """rb_comm_init_all: ONE host process drives two contexts (what a Rust host with a context per GPU does); the J
    partials of the two P-shards are completed by rb_allreduce_sum inside rb_comm_group_start / _end."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, ROOT)
    import ctypes as C
    from oracle.api import Oracle
    from rest_tensors_b200.device import Context, ShardedRI
    from rest_tensors_b200._lib import lib, check
    o = Oracle(); o.load_openblas()
    nb, naux, no = 48, 70, 5
    ctxs = [Context(0), Context(1)]
    hs = (C.c_void_p * 2)(ctxs[0].h, ctxs[1].h)
    check(lib.rb_comm_init_all(hs, 2), "rb_comm_init_all")
    assert [c.comm_world for c in ctxs] == [2, 2] and [c.comm_rank for c in ctxs] == [0, 1]
    c = o.fill_linear(nb * nb, 3, scale=nb ** -0.5)
    cm = c.reshape((nb, nb), order="F")
    dm = np.ascontiguousarray((2.0 * cm[:, :no] @ cm[:, :no].T).reshape(-1, order="F"))
    js, ds = [], []
    for r in range(2):
        with torch.cuda.device(r):
            ctxs[r].bind_stream()
            sh = ShardedRI(ctxs[r], nb, naux, r, 2, comm=False).fill_synthetic()   # no torch.distributed anywhere
            d = sh.dp(torch.from_numpy(dm).to(f"cuda:{r}"))
            js.append(sh.j(d, reduce=False)); ds.append(d[: sh.nx])
    check(lib.rb_comm_group_start(), "rb_comm_group_start")
    for r in range(2):
        ctxs[r].allreduce_sum(js[r])
    check(lib.rb_comm_group_end(), "rb_comm_group_end")
    fulls = [ctxs[r].empty(naux) for r in range(2)]
    check(lib.rb_comm_group_start(), "rb_comm_group_start")
    for r in range(2):
        with torch.cuda.device(r):
            ctxs[r].allgather_shards(ds[r], naux, fulls[r])
    check(lib.rb_comm_group_end(), "rb_comm_group_end")
    for r in range(2):
        torch.cuda.synchronize(r)
    ri = o.fill_ri3ao_symm(nb, 0, naux)
    d_ref = o.ri_dp(ri, dm, nb, naux)
    j_ref = o.ri_j(ri, d_ref, nb, naux)
    for r in range(2):
        assert float(np.max(np.abs(js[r].cpu().numpy() - j_ref)) / np.max(np.abs(j_ref))) <= 1e-10
        assert float(np.max(np.abs(fulls[r].cpu().numpy() - d_ref)) / np.max(np.abs(d_ref))) <= 1e-10
    assert torch.equal(js[0].cpu(), js[1].cpu())
    for c_ in ctxs:
        c_.comm_destroy()
