"""torchrun probe: the P-sharded RPA-type consumer (out[P_local, all Q]) with the peer-memory pipeline (copy-engine pulls
over NVLink overlapped with the DMMA GEMMs) vs NCCL all-gather + GEMMs.  Weak scaling: 1700 slabs (config C) per rank."""
import json, os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
from rest_tensors_b200.device import Context, ShardedRI  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{local}"))
nb, nx, no = 600, int(os.environ.get("PROBE_NX", "1700")), 60
nv = nb - no
ctx = Context(local)
sh = ShardedRI(ctx, nb, nx * world, rank, world).fill_synthetic()
c = ctx.empty(nb * nb); ctx.fill_linear(c, nb * nb, 3, 0, nb ** -0.5)
mo = sh.ao2mo(c[: nb * no], no, c[nb * no:], nv)            # this rank's rows of the occ-vir ri3mo [nx, no, nv]
w = ctx.empty(no * nv); ctx.fill_linear(w, no * nv, 9, 0, 1.0)
out = ctx.empty(sh.nx * sh.naux)
box = (0, no, 0, nv)
res = {}
for name, kw in [("p2p_pipeline", dict(exchange="p2p")), ("nccl_allgather", dict(exchange="allgather"))]:
    for weights in (None, w):
        ts = []
        for it in range(4):
            dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); sh.mo_pq(mo, no, nv, box, weights, out=out, **kw); e1.record()
            torch.cuda.synchronize()
            if it:
                ts.append(e0.elapsed_time(e1))
        t = torch.tensor([min(ts)], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        flop = 2.0 * sh.nx * sh.naux * no * nv * world            # all ranks, full (non-symmetric) block rows
        res[f"{name}{'_weighted' if weights is not None else ''}"] = {"ms": float(t.item()), "tflops_all_ranks": flop / float(t.item()) / 1e9}
    if name == "p2p_pipeline":
        ref = out.clone()
    else:
        res["max_rel_diff_vs_p2p"] = float((out - ref).abs().max() / ref.abs().max())
if rank == 0:
    res.update(world=world, rows_per_rank=sh.nx, pairs=no * nv, panel_MB=sh.nx * no * nv * 8 / 1e6,
               note="host barrier + sync inside the timed region (both variants); best of 3, max over ranks")
    print(json.dumps(res))
    json.dump(res, open(f"gpurun_out/mo_pq_dist_n{world}.json", "w"), indent=1)
dist.destroy_process_group()
