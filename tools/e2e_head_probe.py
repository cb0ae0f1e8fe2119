"""Streaming pass at config C with different head-chunk schedules (REST_B200_HEAD) and chunk sizes (REST_B200_PC), three modes."""
import ctypes as C, os, sys, time
sys.path.insert(0, ".")
import torch
from rest_tensors_b200 import lib
from rest_tensors_b200._lib import check
nb, nx, no = 600, 1700, 60
n2 = nb * nb
mk = lambda n: torch.empty(n, dtype=torch.float64, pin_memory=True)
ri = mk(nx * n2).uniform_(-1, 1); mo = mk(nx * n2)
c = mk(n2).uniform_(-0.04, 0.04); dm = mk(n2).uniform_(-1, 1)
ct = c[: nb * no].clone().pin_memory()
d, j, k = mk(nx), mk(n2), mk(n2)
P = lambda t: C.c_void_p(t.data_ptr())
modes = {
    "full": lambda: lib.rb_host_ri_ao2mo_jk(P(c), nb, P(c), nb, P(ri), P(mo), nb, nx, P(dm), P(ct), no, P(d), P(j), P(k)),
    "upper": lambda: lib.rb_host_ri_ao2mo_jk_upper(P(c), nb, P(ri), P(mo), nb, nx, P(dm), P(ct), no, P(d), P(j), P(k)),
    "symm": lambda: lib.rb_host_ri_ao2mo_jk_symm(P(c), nb, P(ri), P(mo), nb, nx, P(dm), P(ct), no, P(d), P(j), P(k)),
}
for fn in modes.values():
    check(fn(), "warm")
print(f"{'HEAD':>12} {'PC':>5} | " + " ".join(f"{m:>8}" for m in modes))
import json
sched = json.loads(os.environ.get("PROBE_SCHED", "null")) or [("", 256), ("", 192)]
for head, pc in sched:
    os.environ["REST_B200_HEAD"] = head; os.environ["REST_B200_PC"] = str(pc)
    row = []
    for name, fn in modes.items():
        ts = []
        for rep in range(4):
            t0 = time.perf_counter(); check(fn(), name); ts.append(time.perf_counter() - t0)
        row.append(min(ts) * 1e3)
    print(f"{head!r:>12} {pc:>5} | " + " ".join(f"{t:8.1f}" for t in row), flush=True)
