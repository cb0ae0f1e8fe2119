"""Peer-memory operands over NVLink (one process, two devices): (a) pull bandwidth of our strided copy kernel reading the
peer's HBM, (b) the 'N','T' GEMM of the RPA-type consumer with its B operand local vs in the peer's HBM (TMA loads over
NVLink), (c) the same with the weighted pull (scaled local copy of the remote panel, then a local GEMM)."""
import json, sys
import torch
sys.path.insert(0, ".")
from rest_tensors_b200.device import Context  # noqa: E402
from rest_tensors_b200._lib import lib, check  # noqa: E402


def best_ms(fn, dev, reps=4, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize(dev)
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(dev)
        ts.append(e0.elapsed_time(e1))
    return min(ts)


assert torch.cuda.device_count() >= 2, "needs 2 GPUs"
torch.cuda.set_device(0)
c0, c1 = Context(0), Context(1)
check(lib.rb_peer_enable(c0.h, 1), "rb_peer_enable")
out = {}
# (a) pull bandwidth
n = 1 << 27  # 1 GiB
with torch.cuda.device(1):
    src = c1.empty(n); c1.fill_linear(src, n, 1, 0, 1.0); torch.cuda.synchronize(1)
dst = c0.empty(n); loc = c0.empty(n); c0.fill_linear(loc, n, 1, 0, 1.0)
rows = 1 << 13
ms_r = best_ms(lambda: c0.copy_mm(rows, n // rows, src, rows, n // rows, 0, 0, dst, rows, n // rows, 0, 0), 0)
ms_l = best_ms(lambda: c0.copy_mm(rows, n // rows, loc, rows, n // rows, 0, 0, dst, rows, n // rows, 0, 0), 0)
out["copy_1GiB"] = {"remote_read_GBs": n * 8 / ms_r / 1e6, "local_read_GBs": n * 8 / ms_l / 1e6,
                    "bitwise_equal": bool(torch.equal(dst, loc))}
del src, dst, loc
# (b), (c) the consumer's GEMM: out[P, Q] = sum_c w_c A[P, c] B[Q, c]
for (m, nq, k) in [(213, 213, 32400), (850, 850, 32400), (600, 600, 291600 // 8)]:
    ld = m + (m & 1)
    a = c0.empty(ld * k); c0.fill_linear(a, ld * k, 2, 0, 1.0)
    bl = c0.empty(ld * k); c0.fill_linear(bl, ld * k, 3, 0, 1.0)
    with torch.cuda.device(1):
        br = c1.empty(ld * k); c1.fill_linear(br, ld * k, 3, 0, 1.0); torch.cuda.synchronize(1)
    w = c0.empty(k); c0.fill_linear(w, k, 4, 0, 1.0)
    o1 = c0.empty(m * nq); o2 = c0.empty(m * nq)
    row = {}
    for name, b in [("local", bl), ("remote", br)]:
        o = o1 if name == "local" else o2
        ms = best_ms(lambda: c0.ri_mo_pq(a, ld, m, b, ld, nq, 1, k, (0, 1, 0, k), None, 0.0, o, m), 0)
        msw = best_ms(lambda: c0.ri_mo_pq(a, ld, m, b, ld, nq, 1, k, (0, 1, 0, k), w, 0.0, o, m), 0)
        row[name] = {"tma_direct_ms": ms, "tma_direct_tflops": 2.0 * m * nq * k / ms / 1e9,
                     "weighted_pull_ms": msw, "weighted_pull_tflops": 2.0 * m * nq * k / msw / 1e9}
    c0.ri_mo_pq(a, ld, m, bl, ld, nq, 1, k, (0, 1, 0, k), None, 0.0, o1, m)
    c0.ri_mo_pq(a, ld, m, br, ld, nq, 1, k, (0, 1, 0, k), None, 0.0, o2, m)
    torch.cuda.synchronize(0)
    row["bitwise_equal"] = bool(torch.equal(o1, o2))
    row["panel_MB"] = ld * k * 8 / 1e6
    out[f"m{m}_n{nq}_k{k}"] = row
    del a, bl, br, w, o1, o2
for k_, v in out.items():
    print(k_, json.dumps(v))
json.dump(out, open("gpurun_out/peer_probe.json", "w"), indent=1)
