"""torchrun probe: special_dgemm_f_01 contracted over the sharded auxiliary index (T <- T B, B = [naux, naux]) with the
peer-memory pipeline; weak scaling, config-C-sized shard per rank (nb = 600, 1700 slabs = 4.9 GB)."""
import json, os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
from rest_tensors_b200.device import Context, ShardedRI  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{local}"))
nb, nx = 600, int(os.environ.get("PROBE_NX", "1700"))
ctx = Context(local)
sh = ShardedRI(ctx, nb, nx * world, rank, world).fill_synthetic()
naux = sh.naux
b = ctx.empty(naux * naux); ctx.fill_linear(b, naux * naux, 8, 0, naux ** -0.5)
ts = []
for it in range(3):
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sh.special_dgemm_p(b, 1.0, 0.0); e1.record()
    torch.cuda.synchronize()
    if it:
        ts.append(e0.elapsed_time(e1))
t = torch.tensor([min(ts)], dtype=torch.float64, device=f"cuda:{local}")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    flop = 2.0 * nb * nb * naux * naux
    pulled = (world - 1) * nb * nb * sh.nx * 8
    res = {"world": world, "nb": nb, "naux": naux, "ms": float(t.item()), "tflops_all_ranks": flop / float(t.item()) / 1e9,
           "tflops_per_gpu": flop / float(t.item()) / 1e9 / world, "pulled_GB_per_rank": pulled / 1e9,
           "pull_time_at_770GBs_ms": pulled / 770e9 * 1e3,
           "note": "includes IPC export/open, 2 host barriers and the output allocation; best of 2, max over ranks"}
    print(json.dumps(res))
    json.dump(res, open(f"gpurun_out/special_dgemm_dist_n{world}.json", "w"), indent=1)
dist.destroy_process_group()
