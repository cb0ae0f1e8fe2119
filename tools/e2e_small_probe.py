"""Streaming pass (pinned buffers) at configs A and B for several chunk sizes (REST_B200_PC): ms per call, best of 6."""
import ctypes as C, os, sys, time
sys.path.insert(0, ".")
import torch
from rest_tensors_b200 import lib
from rest_tensors_b200._lib import check
mk = lambda n: torch.empty(n, dtype=torch.float64, pin_memory=True)
P = lambda t: C.c_void_p(t.data_ptr())
for nb, nx, no in [(100, 400, 20), (264, 720, 21)]:
    n2 = nb * nb
    ri = mk(nx * n2).uniform_(-1, 1); mo = mk(nx * n2)
    c = mk(n2).uniform_(-0.04, 0.04); dm = mk(n2).uniform_(-1, 1)
    ct = c[: nb * no].clone().pin_memory()
    d, j, k = mk(nx), mk(n2), mk(n2)
    fn = lambda: lib.rb_host_ri_ao2mo_jk(P(c), nb, P(c), nb, P(ri), P(mo), nb, nx, P(dm), P(ct), no, P(d), P(j), P(k))
    check(fn(), "warm")
    out = []
    for pc in (256, 192, 128, 96, 64, 48, 32, 256):
        os.environ["REST_B200_PC"] = str(pc)
        ts = []
        for _ in range(8):
            t0 = time.perf_counter(); check(fn(), "call"); ts.append(time.perf_counter() - t0)
        out.append(f"pc={pc}: {min(ts[2:]) * 1e3:.3f}")
    print(f"nb={nb} nx={nx}: " + "  ".join(out), flush=True)
