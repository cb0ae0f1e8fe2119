"""Per-chunk timeline (REST_B200_TRACE=1) of the streaming pass at config C in its three output / input modes:
full (rb_host_ri_ao2mo_jk), upper pairs out (..._upper), symmetric slabs in + upper pairs out (..._symm)."""
import ctypes as C, os, sys, time
os.environ["REST_B200_TRACE"] = "1"
sys.path.insert(0, ".")
import torch
from rest_tensors_b200 import lib
from rest_tensors_b200._lib import check
nb, nx, no = 600, 1700, 60
n2 = nb * nb
npair = nb * (nb + 1) // 2
mk = lambda n: torch.empty(n, dtype=torch.float64, pin_memory=True)
ri = mk(nx * n2).uniform_(-1, 1)
mo = mk(nx * n2)
c = mk(n2).uniform_(-0.04, 0.04); dm = mk(n2).uniform_(-1, 1)
ct = c[: nb * no].clone().pin_memory()
d, j, k = mk(nx), mk(n2), mk(n2)
P = lambda t: C.c_void_p(t.data_ptr())
modes = {
    "full": lambda: lib.rb_host_ri_ao2mo_jk(P(c), nb, P(c), nb, P(ri), P(mo), nb, nx, P(dm), P(ct), no, P(d), P(j), P(k)),
    "upper": lambda: lib.rb_host_ri_ao2mo_jk_upper(P(c), nb, P(ri), P(mo), nb, nx, P(dm), P(ct), no, P(d), P(j), P(k)),
    "symm": lambda: lib.rb_host_ri_ao2mo_jk_symm(P(c), nb, P(ri), P(mo), nb, nx, P(dm), P(ct), no, P(d), P(j), P(k)),
    "symm, J/K only": lambda: lib.rb_host_ri_ao2mo_jk_symm(P(c), nb, P(ri), None, nb, nx, P(dm), P(ct), no, P(d), P(j), P(k)),
    "full, J/K only": lambda: lib.rb_host_ri_ao2mo_jk(P(c), nb, P(c), nb, P(ri), None, nb, nx, P(dm), P(ct), no, P(d), P(j), P(k)),
}
for name, fn in modes.items():
    print("==", name, flush=True)
    ts = []
    for rep in range(3):
        sys.stderr.flush()
        t0 = time.perf_counter(); check(fn(), name); ts.append(time.perf_counter() - t0)
    print(f"== {name}: {min(ts[1:]) * 1e3:.1f} ms per call (best of 2 after warm-up)", flush=True)
