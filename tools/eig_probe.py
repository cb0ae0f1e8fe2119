"""Eigen-solver timings: one-sided Jacobi on the GPU (device-resident rb_dsyev / rb_dspgv / rb_matrix_power) next to the
reference's LAPACK calls (dsyev / dspgvx / _power of the oracle's OpenBLAS on the box's host cores)."""
import json, sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from rest_tensors_b200.device import Context  # noqa: E402
from oracle.api import Oracle  # noqa: E402

ctx = Context(0)
o = Oracle(); o.load_openblas()
out = {"host_threads": o.blas_threads()}
for n in (264, 600, 1800):
    a = o.fill_linear(n * n, 71).reshape((n, n), order="F"); a = a + a.T
    b = o.fill_linear(n * n, 76).reshape((n, n), order="F"); b = b @ b.T / n + np.eye(n)
    flat = lambda m: np.ascontiguousarray(m.reshape(-1, order="F"))  # noqa: E731
    pack = lambda m: np.ascontiguousarray(np.concatenate([m[: j + 1, j] for j in range(n)]))  # noqa: E731
    ad = torch.from_numpy(flat(a)).cuda(); apd = torch.from_numpy(pack(a)).cuda(); bpd = torch.from_numpy(pack(b)).cuda()
    bd = torch.from_numpy(flat(b)).cuda()
    w = ctx.empty(n); z = ctx.empty(n * n); x = ctx.empty(n * n)
    row = {}
    for name, fn in [("dsyev", lambda: ctx.dsyev("V", "L", n, ad, n, w, z, n)),
                     ("dspgv_half", lambda: ctx.dspgv(n, apd, bpd, n // 2, w, z, n)),
                     ("power_-0.5", lambda: ctx.matrix_power(n, bd, n, -0.5, 1e-10, x, n))]:
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
        row[name + "_gpu_ms"] = min(ts)
    t0 = time.perf_counter(); zr, wr = o.dsyev(flat(a), n); row["dsyev_lapack_ms"] = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter(); o.dspgvx(pack(a), pack(b), n, n // 2); row["dspgvx_half_lapack_ms"] = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter(); o.power(flat(b), n, -0.5, 1e-10); row["power_lapack_ms"] = (time.perf_counter() - t0) * 1e3
    ctx.dsyev("V", "L", n, ad, n, w, z, n)
    zz = z.cpu().numpy().reshape((n, n), order="F"); ww = w.cpu().numpy()
    row["eigenvalue_err_rel"] = float(np.max(np.abs(ww - wr)) / np.max(np.abs(wr)))
    row["residual_rel"] = float(np.max(np.abs(a @ zz - zz * ww)) / np.max(np.abs(wr)))
    row["orthogonality"] = float(np.max(np.abs(zz.T @ zz - np.eye(n))))
    out[f"n{n}"] = row
    print(n, json.dumps(row))
json.dump(out, open("gpurun_out/eig_probe.json", "w"), indent=1)
