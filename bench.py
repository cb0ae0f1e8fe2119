#!/usr/bin/env python
"""bench.py -- RI ao2mo + JK build FP64 GFLOP/s on B200 (BASELINE.json metric), next to the reference CPU path.

A "step" is one pass of the hot path over one rank's P-shard of a synthetic RI tensor:
    ao2mo (square C, reference semantics: 4*nb^3*nx flop)  +  d_P  +  J  +  K  (+ ONE all-reduce each for J, K when N > 1)
Workload at N=1: BASELINE config C (nb=600, naux=1700, nocc=60; ri3ao 4.9 GB) -- the largest configuration of
BASELINE.json:configs that fits one GPU together with its square ri3mo and host staging (config D's square ri3mo
is 124 GB + 124 GB).  N > 1 is weak scaling: every rank holds `nx` slabs (global naux = nx * N, P-sharded exactly
like shard_range()).  `--config D` selects the north-star shard (nb=1800, 600 slabs/rank = naux 4800 on 8 GPUs).

    python bench.py --gpus N --steps K --warmup W              # our arm (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...    # the reference algorithm on the host cores (rank 0 only)

`value` is device-timed (CUDA events, barrier + synchronize on both sides, max over ranks) with inputs resident in
HBM; `e2e` goes through the host-pointer C ABI (rb_host_ri_ao2mo_jk: pinned host ri3ao in, host ri3mo/J/K out, H2D
and D2H inside the timed region).  Inputs (4.9 GB) are larger than L2 (126 MB), so no explicit L2 flush is needed.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CONFIGS = {  # name: (nb, slabs per rank, nocc, description)
    "A": (100, 400, 20, "A: bench_tensors.rs scale nb=100 naux=400 nocc=20"),
    "B": (264, 720, 21, "B: benzene/def2-TZVP-sized nb=264 naux=720 nocc=21"),
    "C": (600, 1700, 60, "C: C20/cc-pVTZ-sized nb=600 naux=1700 nocc=60"),
    "D": (1800, 600, 180, "D: C60/cc-pVTZ-sized shard nb=1800 naux=4800/8 nocc=180"),
}
METRIC = "RI ao2mo + JK build FP64 GFLOP/s"
UNIT = "GFLOP/s"


_REAL_STDOUT = None


def quiet_stdout():
    """Libraries (NCCL's version banner, OpenBLAS warnings) write to fd 1; the contract is ONE JSON line on stdout.
    Point fd 1 at stderr for the duration of the run and keep the real stdout for emit_json_line()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit_json_line(obj):
    sys.stdout.flush()
    data = (json.dumps(obj) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def flops(nb, nx, no):
    """algorithmic flop of one step over nx slabs"""
    return {
        "ao2mo": 4.0 * nb ** 3 * nx,
        "k": (2.0 * nb * nb * no + nb * (nb + 1.0) * no) * nx,
        "dp": 2.0 * nb * nb * nx,
        "j": 2.0 * nb * nb * nx,
    }


# --------------------------------------------------------------------------------------------------------------
# reference CPU path (oracle port + OpenBLAS) -- used by --impl reference and by the cpu_baseline leg
# --------------------------------------------------------------------------------------------------------------
class CpuPath:
    def __init__(self, nb, no):
        import numpy as np
        from oracle.api import Oracle
        self.np = np
        self.o = Oracle()
        self.cores = os.cpu_count() or 1
        try:
            self.cores = len(os.sched_getaffinity(0))
        except Exception:
            pass
        self.have_blas = self.o.load_openblas(threads=self.cores)
        self.threads = self.o.blas_threads() if self.have_blas else 1
        self.nb, self.no = nb, no
        c = self.o.fill_linear(nb * nb, 3, scale=nb ** -0.5)
        cm = c.reshape((nb, nb), order="F")
        self.c = c
        self.dm = np.ascontiguousarray((2.0 * cm[:, :no] @ cm[:, :no].T).reshape(-1, order="F"))
        self.ct = np.ascontiguousarray((cm[:, :no] * np.sqrt(2.0)).reshape(-1, order="F"))
        self.ri = None
        self.ns = 0

    def set_sample(self, slabs):
        self.ns = int(slabs)
        self.ri = self.o.fill_ri3ao_symm(self.nb, 0, self.ns)

    def step(self):
        """reference algorithm on the sample: ri_ao2mo_f (restmatr.f90:158-194) + d_P/J (dgemv) + K (dgemm+dsyrk per slab)"""
        o, nb, ns = self.o, self.nb, self.ns
        mo = o.ri_ao2mo_f(self.c, self.ri, nb, nb, ns)
        d = o.ri_dp(self.ri, self.dm, nb, ns)
        j = o.ri_j(self.ri, d, nb, ns)
        k = o.ri_k(self.ri, self.ct, nb, self.no, ns)
        return mo, d, j, k

    def calibrate(self, target_s, max_slabs):
        """pick a sample size whose step takes about target_s seconds"""
        probe = max(1, min(4, max_slabs))
        self.set_sample(probe)
        self.step()
        t0 = time.perf_counter(); self.step(); dt = time.perf_counter() - t0
        per_slab = max(dt / probe, 1e-6)
        slabs = int(max(1, min(max_slabs, target_s / per_slab)))
        self.set_sample(slabs)
        return slabs

    def describe(self, slabs, nx):
        return (f"{slabs} of {nx} slabs (nb={self.nb}, nocc={self.no}); ri_ao2mo_f loop (dgemm NN + dgemm TN + strided "
                f"scatter per slab) + dgemv T/N + per-slab dgemm/dsyrk; {self.o.blas_config()}")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    quiet_stdout()
    nb, nx, no, desc = CONFIGS[args.config]
    cpu = CpuPath(nb, no)
    slabs = cpu.calibrate(target_s=3.0, max_slabs=nx)
    for _ in range(args.warmup):
        cpu.step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu.step()
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    f = flops(nb, slabs, no)
    value = sum(f.values()) / dt / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "impl": "reference",
        "config": {"workload": desc, "nb": nb, "slabs_per_rank": nx, "nocc": no,
                   "note": "reference algorithm (oracle port of restmatr.f90 + OpenBLAS) on the host cores; each step is a "
                           "bounded sample of the per-rank workload, throughput is per-slab so it scales linearly"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu.threads, **host_cpu_model(), "kind": "port",
                         "sample": cpu.describe(slabs, nx)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_json_line(line)
    return 0


# --------------------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md "clocks line")
# --------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 and len(r) >= 8] or [r for _, r in self.rows if len(r) >= 8]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[1]) for r in rows)
        reasons = []
        for idx, name in [(4, "hw_slowdown"), (5, "hw_thermal_slowdown"), (6, "sw_thermal_slowdown"), (7, "sw_power_cap")]:
            if any(r[idx].lower().startswith("active") for r in rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "reasons": reasons,
                "power_w_max": max(float(r[3]) for r in rows), "samples": len(rows)}


def host_cpu_model():
    """CPU model string and logical CPU count of the box (SURVEY 8d asks for both next to the CPU baseline)"""
    model = None
    try:
        for line in open("/proc/cpuinfo"):
            if line.lower().startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except Exception:
        pass
    return {"cpu_model": model, "nproc": os.cpu_count()}


def host_mem_available_bytes():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) * 1024
    except Exception:
        pass
    return None


def run_e2e(args, torch, dist, lib, check, all_reduce_sum, world, dev, nb, nx, no, n2, sh, c, dm, ct, mo, k, total_flop, barrier):
    """Same step through rb_host_ri_ao2mo_jk with pinned HOST buffers: every rank streams its whole shard up and its
    ri3mo rows down.  Returns the `e2e` object; never raises (a host that cannot pin 2 x shard bytes per rank gets a
    reduced-slab measurement, clearly labelled)."""
    import ctypes as C
    if args.no_e2e:
        return {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "skipped": "--no-e2e"}
    # Host placement: pin this rank's host buffers next to its GPU's PCIe root (8 ranks streaming through the far
    # socket share one inter-socket link otherwise).  Restored afterwards so the CPU baseline sees every core.
    saved_affinity = os.sched_getaffinity(0)
    node = C.c_int(-1)
    if not args.no_numa:
        check(lib.rb_bind_host_to_device_numa(int(dev.split(":")[1]), C.byref(node)), "rb_bind_host_to_device_numa")
    try:
        out = _run_e2e_bound(args, torch, dist, lib, check, all_reduce_sum, world, dev, nb, nx, no, n2, sh, c, dm, ct, mo, k,
                             total_flop, barrier)
    finally:
        os.sched_setaffinity(0, saved_affinity)
    out["host_numa_node"] = int(node.value)
    return out


def _run_e2e_bound(args, torch, dist, lib, check, all_reduce_sum, world, dev, nb, nx, no, n2, sh, c, dm, ct, mo, k, total_flop,
                   barrier):
    import ctypes as C
    slabs = nx
    avail = host_mem_available_bytes()
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    need = 2 * nx * n2 * 8 * local_world
    note = None
    if avail is not None and need > 0.6 * avail:
        slabs = max(64, int(nx * 0.6 * avail / need) // 64 * 64)
        note = f"host RAM allows pinning only {slabs} of {nx} slabs per rank; e2e measured on that sub-shard"
    alloc_err = None
    try:
        ri_h = torch.empty(slabs * n2, dtype=torch.float64, pin_memory=True)
        ri_h.copy_(sh.data[: slabs * n2])
        mo_h = torch.empty(slabs * n2, dtype=torch.float64, pin_memory=True)
        c_h, dm_h, ct_h = c.cpu().pin_memory(), dm.cpu().pin_memory(), ct.cpu().pin_memory()
        d_h = torch.empty(slabs, dtype=torch.float64, pin_memory=True)
        j_h = torch.empty(n2, dtype=torch.float64, pin_memory=True)
        k_h = torch.empty(n2, dtype=torch.float64, pin_memory=True)
        torch.cuda.synchronize()
    except Exception as exc:  # noqa: BLE001
        alloc_err = f"{type(exc).__name__}: {exc}"[:300]
    flag = torch.tensor([0.0 if alloc_err else 1.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)   # all ranks take the same branch
    if float(flag.item()) < 1.0:
        return {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "error": alloc_err or "another rank could not pin its host buffers"}
    try:
        P = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731

        def e2e_step():
            check(lib.rb_host_ri_ao2mo_jk(P(c_h), nb, P(c_h), nb, P(ri_h), P(mo_h), nb, slabs, P(dm_h), P(ct_h), no, P(d_h),
                                          P(j_h), P(k_h)), "rb_host_ri_ao2mo_jk")
            if world > 1:  # complete J and K across ranks (host results -> NVLink all-reduce -> host)
                jk = torch.cat([j_h, k_h]).to(dev, non_blocking=True)
                all_reduce_sum(jk, world)
                jk_h = jk.cpu()
                j_h.copy_(jk_h[:n2]); k_h.copy_(jk_h[n2:])

        steps = max(2, min(args.steps, 5))
        e2e_step()
        # parity spot check against the device-resident results of the same inputs (before J/K get all-reduced again)
        ok = bool(torch.allclose(mo_h[: 4096], mo.view(-1)[: 4096].cpu(), rtol=1e-12, atol=1e-14)) if slabs == nx else None
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            e2e_step()
        barrier()
        dt = torch.tensor([(time.perf_counter() - t0) / steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        flop = total_flop * slabs / nx
        out = {"value": flop / float(dt.item()) / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": (slabs * n2 + 2 * n2 + nb * no) * 8, "d2h_bytes_per_step": (slabs * n2 + 2 * n2 + slabs) * 8,
               "ms_per_step": float(dt.item()) * 1e3,
               "api": "rb_host_ri_ao2mo_jk (host-pointer C ABI; pinned host buffers; 3-stream H2D|compute|D2H pipeline)",
               "steps": steps, "matches_device_path": ok}
        if note:
            out["note"] = note
        return out
    except Exception as exc:  # noqa: BLE001
        return {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "error": f"{type(exc).__name__}: {exc}"[:300]}


# --------------------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------------------
def run_ours(args):
    quiet_stdout()
    import ctypes as C
    import numpy as np
    import torch
    import torch.distributed as dist
    from rest_tensors_b200 import lib
    from rest_tensors_b200._lib import check
    from rest_tensors_b200.device import Context, ShardedRI, all_reduce_sum

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs CUDA devices (no CPU fallback)"
    torch.cuda.set_device(local)
    os.environ["REST_B200_DEVICE"] = str(local)   # the host-pointer C ABI (e2e leg) runs on this rank's GPU
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{local}"))
    dev = f"cuda:{local}"
    nb, nx, no, desc = CONFIGS[args.config]
    naux = nx * world
    ctx = Context(local)
    sh = ShardedRI(ctx, nb, naux, rank, world).fill_synthetic()
    assert sh.nx == nx
    n2 = nb * nb
    c = ctx.empty(n2); ctx.fill_linear(c, n2, 3, 0, nb ** -0.5)
    ct = ctx.empty(nb * no); ctx.fill_linear(ct, nb * no, 3, 0, nb ** -0.5); ctx.self_multiple(ct, 2.0 ** 0.5, nb * no)
    dm = ctx.empty(n2)
    ctx.dgemm("N", "T", nb, nb, no, 2.0, c, nb, c, nb, 0.0, dm, nb)       # D = 2 C_occ C_occ^T (our own GEMM)
    mo = ctx.empty(nx * n2); d = ctx.empty(nx); j = ctx.empty(n2); k = ctx.empty(n2)
    f = flops(nb, nx, no)
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    seg = {"ao2mo": [], "dp": [], "j": [], "k": []}

    def step(record=False):
        marks = [ev() for _ in range(5)] if record else None
        if record: marks[0].record()
        sh.ao2mo(c, nb, c, nb, out=mo)
        if record: marks[1].record()
        sh.dp(dm, out=d)
        if record: marks[2].record()
        sh.j(d, out=j, reduce=True)
        if record: marks[3].record()
        sh.k(ct, no, out=k, reduce=True)
        if record: marks[4].record()
        return marks

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # roofline denominators, measured live before the timed region
    dmma_peak = max(ctx.fp64_peak_probe(0, 100000)[0] for _ in range(2))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak, hbm_src = (peaks["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (of measured)") if "hbm_gbs" in peaks else \
        (6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)")

    for _ in range(max(args.warmup, 3) if args.warmup >= 0 else 3):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.2)
    launches0 = ctx.launches
    barrier()
    t_wall0 = time.perf_counter()
    e0, e1 = ev(), ev()
    e0.record()
    all_marks = [step(record=True) for _ in range(args.steps)]
    e1.record()
    barrier()
    t_wall1 = time.perf_counter()
    launches = ctx.launches - launches0
    ms_total = e0.elapsed_time(e1)
    tmax = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_step = float(tmax.item()) / args.steps
    for m in all_marks:
        for i, name in enumerate(["ao2mo", "dp", "j", "k"]):
            seg[name].append(m[i].elapsed_time(m[i + 1]))
    avg = {kname: sum(v) / len(v) for kname, v in seg.items()}
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    total_flop = sum(f.values()) * world
    value = total_flop / (ms_step * 1e-3) / 1e9

    # ---- extra (outside the timed region, not part of `value`): the north star's occ-vir form of ao2mo,
    #      C_occ^T (mu nu|P) C_vir, on the same shard; flop = (2 no nb^2 + 2 no nb nv) nx ----
    nv = nb - no
    ov = ctx.empty(nx * no * nv)
    ov_ms = []
    for it in range(0 if args.no_extras else 4):
        a0, a1 = ev(), ev()
        a0.record(); sh.ao2mo(c[: nb * no], no, c[nb * no:], nv, out=ov); a1.record()
        torch.cuda.synchronize()
        if it:
            ov_ms.append(a0.elapsed_time(a1))
    ov_flop = (2.0 * no * nb * nb + 2.0 * no * nb * nv) * nx
    occ_vir = None if args.no_extras else {
        "ms": min(ov_ms), "tflops_per_gpu": ov_flop / (min(ov_ms) * 1e-3) / 1e12,
        "note": "ao2mo with C_left = C_occ [nb,nocc], C_right = C_vir [nb,nb-nocc]; best of 3, per rank"}
    # ---- extra: the step after ao2mo -- (ia|jb) blocks straight from the occ-vir ri3mo above:
    #      diagonal block pair (i-block == j-block; SYRK, M(M+1)K flop) and an off-diagonal pair (GEMM, 2MNK flop);
    #      the i-block is as many occupied orbitals as keep the [M, M] block under 6 GB ----
    iajb = None
    try:
        if args.no_extras:
            raise RuntimeError("skipped (--no-extras)")
        li = max(1, min(no // 2 if no >= 2 else 1, int(((6 << 30) / 8) ** 0.5) // max(nv, 1)))
        if (6 << 30) / 8 >= float(no * nv) ** 2:
            li = no
        m_blk = li * nv
        g = ctx.empty(m_blk * m_blk)
        res = {}
        pairs = [("diag", (0, li, 0, nv), (0, li, 0, nv), float(m_blk) * (m_blk + 1) * nx)]
        if 2 * li <= no:
            pairs.append(("offdiag", (0, li, 0, nv), (li, li, 0, nv), 2.0 * m_blk * m_blk * nx))
        for name, ba, bb, fl in pairs:
            best = None
            for it in range(3):
                a0, a1 = ev(), ev()
                a0.record(); sh.iajb(ov, no, nv, ba, bb, out=g, reduce=False); a1.record()
                torch.cuda.synchronize()
                if it:
                    best = a0.elapsed_time(a1) if best is None else min(best, a0.elapsed_time(a1))
            res[name] = {"ms": best, "tflops_per_gpu": fl / (best * 1e-3) / 1e12}
        # RPA-type consumer on the same tensor: Pi[P,Q] = sum_ia w_ia R_ia^P R_ia^Q (upper triangle + mirror, np(np+1)K flop)
        wts = ctx.empty(no * nv); ctx.fill_linear(wts, no * nv, 9, 0, 1.0)
        pi = ctx.empty(nx * nx)
        best = None
        for it in range(3):
            a0, a1 = ev(), ev()
            a0.record(); ctx.ri_mo_pq(ov, nx, nx, ov, nx, nx, no, nv, (0, no, 0, nv), wts, 0.0, pi, nx); a1.record()
            torch.cuda.synchronize()
            if it:
                best = a0.elapsed_time(a1) if best is None else min(best, a0.elapsed_time(a1))
        res["mo_pq_weighted"] = {"ms": best, "tflops_per_gpu": float(nx) * (nx + 1) * no * nv / (best * 1e-3) / 1e12,
                                 "m": nx, "k": no * nv}
        del pi, wts
        iajb = {"occ_block": li, "rows": m_blk, "k": nx, **res,
                "note": "rb_ri_iajb on this rank's rows of the occ-vir ri3mo (partial sum; all-reduce not timed); best of 2"}
        del g
    except Exception as exc:  # noqa: BLE001
        iajb = {"error": f"{type(exc).__name__}: {exc}"[:200]}
    del ov

    # ---- e2e through the host-pointer C ABI (pinned host buffers; H2D + D2H inside the timed region) ----
    e2e = run_e2e(args, torch, dist, lib, check, all_reduce_sum, world, dev, nb, nx, no, n2, sh, c, dm, ct, mo, k, total_flop,
                  barrier)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- cpu_baseline: the reference algorithm on this box's host cores, bounded sample (N == 1 only) ----
    cpu_baseline = None
    if world == 1 and not args.no_cpu:
        cpu = CpuPath(nb, no)
        slabs = cpu.calibrate(target_s=4.0, max_slabs=nx)
        cpu.step()                                  # warm-up 1, median of 3
        cts = []
        for _ in range(3):
            t0 = time.perf_counter(); cpu.step(); cts.append(time.perf_counter() - t0)
        cdt = sorted(cts)[1]
        cpu_baseline = {"value": sum(flops(nb, slabs, no).values()) / cdt / 1e9, "unit": UNIT, "cores": cpu.threads,
                        "kind": "port", "sample": cpu.describe(slabs, nx), **host_cpu_model()}
        # the same algorithm on ONE host thread (SURVEY 8d asks for both), on a ~3 s sample
        if cpu.have_blas and cpu.threads > 1:
            all_threads = cpu.threads
            cpu.o.set_threads(1)
            s1 = cpu.calibrate(target_s=3.0, max_slabs=slabs)
            t0 = time.perf_counter(); cpu.step(); c1 = time.perf_counter() - t0
            cpu.o.set_threads(all_threads)
            cpu_baseline["value_single_thread"] = sum(flops(nb, s1, no).values()) / c1 / 1e9
            cpu_baseline["single_thread_sample_slabs"] = s1

    traffic = traffic_dp = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        traffic, traffic_dp = tj.get(args.config), tj.get(args.config + "_dp")
    except Exception:
        pass
    gemm_launch_ms = avg["ao2mo"] / 2.0                      # ao2mo = 2 launches of the TMA+DMMA GEMM kernel
    achieved = (f["ao2mo"] / 2.0) / (gemm_launch_ms * 1e-3) / 1e12
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": desc, "nb": nb, "slabs_per_rank": nx, "naux_global": naux, "nocc": no,
                   "parallelism": f"P-shard x{world}; all-reduce(sum) of J and K only",
                   "ao2mo": "square C (reference ri_ao2mo_f semantics), 4*nb^3*nx flop",
                   "l2": "inputs (ri3ao %.1f GB/rank) larger than L2; no flush needed" % (nx * n2 * 8 / 1e9)},
        "breakdown_ms": {kname: round(v, 4) for kname, v in avg.items()},
        "breakdown_rate": {"ao2mo_tflops": f["ao2mo"] / (avg["ao2mo"] * 1e-3) / 1e12,
                           "k_tflops": f["k"] / (avg["k"] * 1e-3) / 1e12,
                           "dp_gbs": nx * n2 * 8 / (avg["dp"] * 1e-3) / 1e9, "j_gbs": nx * n2 * 8 / (avg["j"] * 1e-3) / 1e9},
        "roofline": {"bound": "tensor", "kernel": "rb_gemm_tma_kernel<A_K=1,B_K=1> (ao2mo GEMMs; DMMA.8x8x4 fed by TMA)",
                     "achieved": achieved, "peak": dmma_peak, "unit": "TFLOP/s", "frac": achieved / dmma_peak,
                     "peak_source": "live register-resident DMMA.8x8x4 probe on this GPU (MEASURED_PEAKS.json holds only "
                                    "bf16/HBM; nominal B200 FP64 tensor 37-40 TFLOP/s)",
                     "flop_per_launch": f["ao2mo"] / 2.0, "launch_ms": gemm_launch_ms, "traffic": traffic},
        "roofline_hbm": {"bound": "hbm", "kernel": "rb_gemv_t_vec_kernel (d_P) / rb_gemv_n_kernel (J)",
                         "achieved": nx * n2 * 8 / (avg["dp"] * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": nx * n2 * 8 / (avg["dp"] * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src,
                         "j_achieved": nx * n2 * 8 / (avg["j"] * 1e-3) / 1e9, "traffic": traffic_dp},
        "e2e": e2e,
        "ao2mo_occ_vir": occ_vir,
        "iajb_occ_vir": iajb,
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if cpu_baseline is not None:
        line["cpu_baseline"] = cpu_baseline
    emit_json_line(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-numa", action="store_true", help="e2e leg: do not bind the rank to its GPU's NUMA node")
    ap.add_argument("--no-extras", action="store_true", help="skip the untimed extras (occ-vir ao2mo, ri3mo consumers): "
                    "the ncu launch list then holds the timed step's kernels only")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-pointer e2e leg (e.g. config D: 2 x 15.5 GB pinned per rank)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
