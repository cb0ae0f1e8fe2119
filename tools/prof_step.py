"""Profiling target: one device-resident step (ao2mo + d_P + J + K) at `nb nx no` (default config A); per-segment CUDA-event times
(back to back, warm) when run plainly, per-kernel times when run under ncu (tools/gpu_batch_r02x.sh)."""
import sys
import torch
sys.path.insert(0, ".")
from rest_tensors_b200.device import Context, ShardedRI  # noqa: E402

nb, nx, no = (100, 400, 20) if len(sys.argv) < 4 else (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]))
ctx = Context(0)
sh = ShardedRI(ctx, nb, nx).fill_synthetic()
n2 = nb * nb
c = ctx.empty(n2); ctx.fill_linear(c, n2, 3, 0, nb ** -0.5)
ct = ctx.empty(nb * no); ctx.fill_linear(ct, nb * no, 3, 0, nb ** -0.5)
dm = ctx.empty(n2); ctx.fill_linear(dm, n2, 4, 0, 1.0 / nb)
mo = ctx.empty(nx * n2); d = ctx.empty(nx); j = ctx.empty(n2); k = ctx.empty(n2)


def step(marks=None):
    if marks: marks[0].record()
    sh.ao2mo(c, nb, c, nb, out=mo)
    if marks: marks[1].record()
    sh.dp(dm, out=d)
    if marks: marks[2].record()
    sh.j(d, out=j)
    if marks: marks[3].record()
    sh.k(ct, no, out=k)
    if marks: marks[4].record()


for _ in range(3):
    step()
torch.cuda.synchronize()
acc = [0.0] * 4
reps = 20
for _ in range(reps):
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    step(marks)
    torch.cuda.synchronize()
    for i in range(4):
        acc[i] += marks[i].elapsed_time(marks[i + 1]) * 1e3 / reps
flop_ao = 4.0 * nb ** 3 * nx
flop_k = (2.0 * nb * nb * no + nb * (nb + 1) * no) * nx
print(f"step nb={nb} nx={nx} no={no}: ao2mo {acc[0]:.1f} us ({flop_ao / acc[0] / 1e6:.2f} TFLOP/s)  dp {acc[1]:.1f} us "
      f"({nx * n2 * 8 / acc[1] / 1e3:.0f} GB/s)  j {acc[2]:.1f} us ({nx * n2 * 8 / acc[2] / 1e3:.0f} GB/s)  k {acc[3]:.1f} us "
      f"({flop_k / acc[3] / 1e6:.2f} TFLOP/s)  sum {sum(acc):.1f} us")
