"""Times the host-pointer streaming call (pinned buffers) on config C for several chunk sizes / output modes."""
import ctypes as C, os, subprocess, sys, time
if len(sys.argv) > 1 and sys.argv[1] == "child":
    os.environ["REST_B200_TRACE"] = "1"
    sys.path.insert(0, ".")
    import torch
    from rest_tensors_b200 import lib
    from rest_tensors_b200._lib import check
    nb, nx, no = 600, 1700, 60
    n2 = nb * nb
    pin = os.environ.get("PAGEABLE") != "1"
    ri = torch.empty(nx * n2, dtype=torch.float64, pin_memory=pin).uniform_(-1, 1)
    mo = torch.empty(nx * n2, dtype=torch.float64, pin_memory=pin)
    if not pin:
        mo.zero_()   # touch the pages
    c = torch.empty(n2, dtype=torch.float64, pin_memory=True).uniform_(-0.04, 0.04)
    dm = torch.empty(n2, dtype=torch.float64, pin_memory=True).uniform_(-1, 1)
    ct = c[: nb * no].clone().pin_memory()
    d = torch.empty(nx, dtype=torch.float64, pin_memory=True); j = torch.empty(n2, dtype=torch.float64, pin_memory=True); k = torch.empty(n2, dtype=torch.float64, pin_memory=True)
    P = lambda t: C.c_void_p(t.data_ptr())
    fn = lambda: lib.rb_host_ri_ao2mo_jk(P(c), nb, P(c), nb, P(ri), P(mo), nb, nx, P(dm), P(ct), no, P(d), P(j), P(k))
    ts = []
    for rep in range(4):
        t0 = time.perf_counter(); check(fn(), "step"); ts.append(time.perf_counter() - t0)
    print(f"  total per call: {min(ts[1:])*1e3:.1f} ms (best of 3)  checksum {float(mo[::100003].sum()):.6e}", flush=True)
else:
    for env in [{}, {"PAGEABLE": "1"}, {"PAGEABLE": "1", "REST_B200_HOST_THREADS": "4"}, {"PAGEABLE": "1", "REST_B200_HOST_THREADS": "16"}, {"PAGEABLE": "1", "REST_B200_BOUNCE": "0"}]:
        print("variant", env, flush=True)
        e = dict(os.environ); e.update(env)
        out = subprocess.run([sys.executable, __file__, "child"], env=e, capture_output=True, text=True)
        print("\n".join((out.stdout + out.stderr).strip().splitlines()[-12:]), flush=True)
