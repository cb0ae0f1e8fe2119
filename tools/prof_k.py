"""Profiling target: the K build (rb_ri_k: batched half-transform GEMM + SYRK + split-K reduce + symmetrize) at config C
(default) or `nb nx no` from the command line; run under ncu (launch list / --set full, see tools/gpu_batch_r02c.sh)."""
import sys
import torch
sys.path.insert(0, ".")
from rest_tensors_b200.device import Context, ShardedRI  # noqa: E402

nb, nx, no = (600, 1700, 60) if len(sys.argv) < 4 else (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]))
ctx = Context(0)
sh = ShardedRI(ctx, nb, nx).fill_synthetic()
ct = ctx.empty(nb * no); ctx.fill_linear(ct, nb * no, 3, 0, nb ** -0.5)
k = ctx.empty(nb * nb)
for _ in range(2):
    sh.k(ct, no, out=k, reduce=False)
torch.cuda.synchronize()
# event-timed (not under ncu this number is the real one)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
torch.cuda.synchronize()
ev[0].record()
for _ in range(5):
    sh.k(ct, no, out=k, reduce=False)
ev[1].record()
torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / 5
flop = (2 * nb * nb * no + nb * (nb + 1) * no) * nx
print(f"K nb={nb} nx={nx} no={no}: {ms:.3f} ms  {flop / ms / 1e9:.2f} TFLOP/s")
