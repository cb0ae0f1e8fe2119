"""dsyev at small n: cluster-resident kernel vs graph-replayed round kernels (REST_B200_EIG_CLUSTER=0 disables the former)."""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from rest_tensors_b200.device import Context  # noqa: E402
from oracle.api import Oracle  # noqa: E402

ctx = Context(0)
o = Oracle()
row = {}
for n in (64, 128, 200, 264, 320):
    a = o.fill_linear(n * n, 71).reshape((n, n), order="F"); a = a + a.T
    ad = torch.from_numpy(np.ascontiguousarray(a.reshape(-1, order="F"))).cuda()
    w = ctx.empty(n); z = ctx.empty(n * n)
    ts = []
    for it in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ctx.dsyev("V", "L", n, ad, n, w, z, n)
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    row[n] = round(min(ts[1:]), 3)
print(os.environ.get("REST_B200_EIG_CLUSTER", "1"), json.dumps(row))
