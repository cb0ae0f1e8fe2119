#!/usr/bin/env python
"""bench.py -- RI ao2mo + JK build FP64 GFLOP/s on B200 (BASELINE.json metric), next to the reference CPU path.

A "step" is one pass of the hot path over one rank's P-shard of a synthetic RI tensor:
    ao2mo (square C, reference semantics: 4*nb^3*nx flop)  +  d_P  +  J  +  K  (+ ONE all-reduce each for J, K when N > 1)
`value` at N=1: BASELINE config C (nb=600, naux=1700, nocc=60; ri3ao 4.9 GB) -- the largest configuration of
BASELINE.json:configs that fits one GPU together with its square ri3mo and host staging.  For N > 1 `value` is weak
scaling of that workload (every rank holds 1700 slabs, global naux = 1700 * N, P-sharded exactly like shard_range()), so
that the line stays comparable across N.  The same JSON line also carries, device-timed with the same rules:
    step_occ_vir  the north star's step: occ-vir ao2mo (C_occ^T (mu nu|P) C_vir) + d_P + J + K on the same shards
    strong_C      (N > 1) config C's 1700 slabs split over the N ranks, incl. both all-reduces (BASELINE: "1-8 B200")
    config_D      (N = 8) the north-star target: nb=1800, naux=4800, nocc=180 P-sharded over 8 GPUs, with its own roofline
    parity        all-reduced J / K, d_P and every rank's ri3mo rows against the CPU oracle on a reduced-slab replica of
                  the same shapes (the oracle is the checker here, never on the timed path)
    small_configs (N = 1) configs A and B: device, e2e pinned, e2e pageable and the reference CPU path
    crossover     (N = 1) per-slab host-pointer BLAS / copy calls from 8 caller threads vs OpenBLAS, and the batched entry point
    e2e_variants  (N = 1) the occ-vir form through the host ABI (10x fewer bytes back) and the resident-shard SCF iteration

    python bench.py --gpus N --steps K --warmup W              # our arm (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...    # the reference algorithm on the host cores (rank 0 only)

`value` is device-timed (CUDA events, barrier + synchronize on both sides, max over ranks) with inputs resident in
HBM; `e2e` goes through the host-pointer C ABI (rb_host_ri_ao2mo_jk: pinned host ri3ao in, host ri3mo/J/K out, H2D
and D2H inside the timed region).  Inputs (4.9 GB) are larger than L2 (126 MB), so no explicit L2 flush is needed; the
small configs (A: 32 MB) flush L2 between timed iterations.  The collectives are librest_b200's own (rb_allreduce_sum:
NCCL bound inside the library, on the compute stream); torch.distributed only carries the rendezvous and host barriers.
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

CONFIGS = {  # name: (nb, slabs per rank in the weak-scaling `value`, nocc, description)
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


def flops(nb, nx, no, square=True):
    """algorithmic flop of one step over nx slabs; square=False: occ-vir ao2mo, cheapest order"""
    return {
        "ao2mo": 4.0 * nb ** 3 * nx if square else (2.0 * no * nb * nb + 2.0 * no * nb * (nb - no)) * nx,
        "k": (2.0 * nb * nb * no + nb * (nb + 1.0) * no) * nx,
        "dp": 2.0 * nb * nb * nx,
        "j": 2.0 * nb * nb * nx,
    }


def config_dict(name, world):
    """The `config` object of the JSON line -- the SAME dict in both arms (the driver compares them key by key)."""
    nb, nx, no, desc = CONFIGS[name]
    return {"workload": desc, "nb": nb, "slabs_per_rank": nx, "naux_global": nx * world, "nocc": no,
            "parallelism": f"P-shard x{world}; all-reduce(sum) of J and K only",
            "ao2mo": "square C (reference ri_ao2mo_f semantics), 4*nb^3*nx flop",
            "l2": "inputs (ri3ao %.1f GB/rank) larger than L2; no flush needed" % (nx * nb * nb * 8 / 1e9)}


# --------------------------------------------------------------------------------------------------------------
# reference CPU path (oracle port + OpenBLAS) -- used by --impl reference, the cpu_baseline leg and the parity checker
# --------------------------------------------------------------------------------------------------------------
class CpuPath:
    def __init__(self, nb, no, threads=None):
        import numpy as np
        from oracle.api import Oracle
        self.np = np
        self.o = Oracle()
        self.cores = os.cpu_count() or 1
        try:
            self.cores = len(os.sched_getaffinity(0))
        except Exception:
            pass
        self.have_blas = self.o.load_openblas(threads=threads or self.cores)
        self.threads = self.o.blas_threads() if self.have_blas else 1
        self.nb, self.no = nb, no
        c = self.o.fill_linear(nb * nb, 3, scale=nb ** -0.5)
        cm = c.reshape((nb, nb), order="F")
        self.c = c
        self.dm = np.ascontiguousarray((2.0 * cm[:, :no] @ cm[:, :no].T).reshape(-1, order="F"))
        self.ct = np.ascontiguousarray((cm[:, :no] * np.sqrt(2.0)).reshape(-1, order="F"))
        self.ri = None
        self.ns = 0

    def set_sample(self, slabs, p_lo=0):
        self.ns = int(slabs)
        self.ri = self.o.fill_ri3ao_symm(self.nb, p_lo, p_lo + self.ns)

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
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
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
        "config": config_dict(args.config, world),
        "note": "reference algorithm (oracle port of restmatr.f90 + OpenBLAS) on the host cores; each step is a bounded "
                "sample of the per-rank workload, throughput is per-slab so it scales linearly",
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


# --------------------------------------------------------------------------------------------------------------
# our arm: one rank's device-resident workload
# --------------------------------------------------------------------------------------------------------------
class Env:
    """per-process plumbing shared by the legs"""

    def __init__(self, torch, dist, ctx, rank, world, local):
        self.torch, self.dist, self.ctx, self.rank, self.world, self.local = torch, dist, ctx, rank, world, local
        self.dev = f"cuda:{local}"

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def ev(self):
        return self.torch.cuda.Event(enable_timing=True)

    def max_over_ranks(self, x):
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


class Workload:
    """ri3ao shard of `naux` global slabs in HBM + replicated C, C~, D + the output buffers of one step"""

    def __init__(self, env, nb, naux, no, square_out=True):
        from rest_tensors_b200.device import ShardedRI
        ctx = env.ctx
        self.env, self.nb, self.naux, self.no = env, nb, naux, no
        self.sh = ShardedRI(ctx, nb, naux, env.rank, env.world).fill_synthetic()
        self.nx = self.sh.nx
        n2 = nb * nb
        self.c = ctx.empty(n2); ctx.fill_linear(self.c, n2, 3, 0, nb ** -0.5)
        self.ct = ctx.empty(nb * no); ctx.fill_linear(self.ct, nb * no, 3, 0, nb ** -0.5)
        ctx.self_multiple(self.ct, 2.0 ** 0.5, nb * no)
        self.dm = ctx.empty(n2)
        ctx.dgemm("N", "T", nb, nb, no, 2.0, self.c, nb, self.c, nb, 0.0, self.dm, nb)   # D = 2 C_occ C_occ^T (our own GEMM)
        self.mo = ctx.empty(max(1, self.nx) * n2) if square_out else None
        self.ov = None
        self.d = ctx.empty(max(1, self.nx)); self.j = ctx.empty(n2); self.k = ctx.empty(n2)

    def step(self, square=True, marks=None):
        sh, nb, no = self.sh, self.nb, self.no
        if marks: marks[0].record()
        if square:
            sh.ao2mo(self.c, nb, self.c, nb, out=self.mo)
        else:
            if self.ov is None:
                self.ov = self.env.ctx.empty(max(1, self.nx) * no * (nb - no))
            sh.ao2mo(self.c[: nb * no], no, self.c[nb * no:], nb - no, out=self.ov)
        if marks: marks[1].record()
        sh.dp(self.dm, out=self.d)
        if marks: marks[2].record()
        sh.j(self.d, out=self.j, reduce=True)
        if marks: marks[3].record()
        sh.k(self.ct, no, out=self.k, reduce=True)
        if marks: marks[4].record()

    def measure(self, steps, warmup, square=True):
        """W untimed steps, then exactly `steps` steps bracketed by barrier + synchronize; CUDA events on the launching
        stream; max over ranks.  Returns ms per step, this rank's per-phase averages and the launches in the timed region."""
        env = self.env
        for _ in range(warmup):
            self.step(square)
        env.barrier()
        launches0 = env.ctx.launches
        e0, e1 = env.ev(), env.ev()
        e0.record()
        all_marks = []
        for _ in range(steps):
            marks = [env.ev() for _ in range(5)]
            self.step(square, marks)
            all_marks.append(marks)
        e1.record()
        env.barrier()
        launches = env.ctx.launches - launches0
        ms_step = env.max_over_ranks(e0.elapsed_time(e1)) / steps
        seg = {}
        for i, name in enumerate(["ao2mo", "dp", "j", "k"]):
            seg[name] = sum(m[i].elapsed_time(m[i + 1]) for m in all_marks) / len(all_marks)
        return ms_step, seg, launches

    def total_flop(self, square=True):
        """whole-job flop of one step (all ranks)"""
        return sum(flops(self.nb, self.naux, self.no, square).values())


def rates(nb, nx, no, seg, square=True):
    f = flops(nb, nx, no, square)
    return {"ao2mo_tflops": f["ao2mo"] / (seg["ao2mo"] * 1e-3) / 1e12, "k_tflops": f["k"] / (seg["k"] * 1e-3) / 1e12,
            "dp_gbs": nx * nb * nb * 8 / (seg["dp"] * 1e-3) / 1e9, "j_gbs": nx * nb * nb * 8 / (seg["j"] * 1e-3) / 1e9}


def allreduce_ms(env, n, reps=20):
    """device time of one rb_allreduce_sum over n doubles (max over ranks)"""
    if env.world == 1:
        return 0.0
    buf = env.ctx.empty(n); buf.zero_()
    for _ in range(3):
        env.ctx.allreduce_sum(buf)
    env.barrier()
    e0, e1 = env.ev(), env.ev()
    e0.record()
    for _ in range(reps):
        env.ctx.allreduce_sum(buf)
    e1.record()
    env.barrier()
    return env.max_over_ranks(e0.elapsed_time(e1)) / reps


def parity_leg(env, nb, no, slabs_per_rank):
    """Oracle parity of the multi-rank path on a reduced-slab replica of the same shapes: naux = slabs_per_rank * N slabs,
    P-sharded like the real run, the same kernels and the same all-reduces.  Every rank checks its own ri3mo rows and d_P
    piece against the oracle run on its own slabs (per-slab work is independent); rank 0 checks the all-reduced J and K
    against the oracle over ALL slabs; all ranks must hold bitwise the same J and K.  Norm-wise relative errors."""
    import numpy as np
    torch = env.torch
    naux = slabs_per_rank * env.world
    w = Workload(env, nb, naux, no)
    w.step(True)
    torch.cuda.synchronize()
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(env.world)))
    cpu = CpuPath(nb, no, threads=max(1, (os.cpu_count() or 1) // local_world))
    o = cpu.o

    def err(x, y):
        den = float(np.max(np.abs(y)))
        return float(np.max(np.abs(np.asarray(x) - y))) / den if den > 0 else 0.0

    ri_local = o.fill_ri3ao_symm(nb, w.sh.p_lo, w.sh.p_hi)
    assert np.array_equal(w.sh.data.cpu().numpy(), ri_local), "device generator != oracle generator"
    e_mo = err(w.mo.cpu().numpy(), o.ri_ao2mo_f(cpu.c, ri_local, nb, nb, w.nx))
    d_ref_local = o.ri_dp(ri_local, cpu.dm, nb, w.nx)
    e_dp = err(w.d.cpu().numpy()[: w.nx], d_ref_local)
    e_mo, e_dp = env.max_over_ranks(e_mo), env.max_over_ranks(e_dp)
    # J / K: identical on every rank?  (sum of |x| and a position-weighted sum, compared as max - min over ranks)
    j, k = w.j.cpu().numpy(), w.k.cpu().numpy()
    sig = [float(np.sum(j)), float(np.sum(k)), float(np.dot(j, np.arange(j.size) % 97)), float(np.dot(k, np.arange(k.size) % 89))]
    same = all(env.max_over_ranks(s) == -env.max_over_ranks(-s) for s in sig)
    d_full = w.sh.gather_dp(w.d[: w.nx]) if env.world > 1 else w.d[: w.nx]
    out = {"shape": {"nb": nb, "nocc": no, "naux": naux, "slabs_per_rank": slabs_per_rank}, "mo_all_ranks": e_mo,
           "d_P_all_ranks": e_dp, "jk_identical_on_all_ranks": bool(same)}
    if env.rank == 0:
        ri = o.fill_ri3ao_symm(nb, 0, naux)
        d_ref = o.ri_dp(ri, cpu.dm, nb, naux)
        out["d_P_gathered"] = err(d_full.cpu().numpy(), d_ref)
        out["J"] = err(j, o.ri_j(ri, d_ref, nb, naux))
        out["K"] = err(k, o.ri_k(ri, cpu.ct, nb, no, naux))
        out["bar"] = 1e-10
        out["ok"] = bool(same and max(out["J"], out["K"], e_mo, e_dp, out["d_P_gathered"]) <= 1e-10)
        out["checker"] = "oracle/rest_oracle.c + OpenBLAS (port of restmatr.f90:158-194 and the dgemv / dgemm+dsyrk composition)"
    env.barrier()
    return out


def l2_flusher(env):
    buf = env.torch.empty(256 << 20, dtype=env.torch.uint8, device=env.dev)
    return lambda: buf.fill_(1)


def e2e_leg(env, lib, check, w, total_flop, steps, pinned=True, no_numa=False):
    """Same step through rb_host_ri_ao2mo_jk with HOST buffers: every rank streams its whole shard up and its ri3mo rows
    down.  pinned=False: plain pageable memory (what a Rust Vec<f64> is).  Never raises."""
    import ctypes as C
    saved_affinity = os.sched_getaffinity(0)
    node = C.c_int(-1)
    if not no_numa:
        check(lib.rb_bind_host_to_device_numa(env.local, C.byref(node)), "rb_bind_host_to_device_numa")
    try:
        out = _e2e_bound(env, lib, check, w, total_flop, steps, pinned)
    finally:
        os.sched_setaffinity(0, saved_affinity)
    out["host_numa_node"] = int(node.value)
    return out


def _e2e_bound(env, lib, check, w, total_flop, steps, pinned):
    import ctypes as C
    torch, dist, world, dev = env.torch, env.dist, env.world, env.dev
    nb, nx, no = w.nb, w.nx, w.no
    n2 = nb * nb
    slabs = nx
    avail = host_mem_available_bytes()
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    need = 2 * nx * n2 * 8 * local_world
    note = None
    if avail is not None and need > 0.6 * avail:
        slabs = max(64, int(nx * 0.6 * avail / need) // 64 * 64)
        note = f"host RAM allows only {slabs} of {nx} slabs per rank; e2e measured on that sub-shard"
    alloc_err = None
    try:
        mk = (lambda n: torch.empty(n, dtype=torch.float64, pin_memory=True)) if pinned else \
             (lambda n: torch.empty(n, dtype=torch.float64))
        ri_h = mk(slabs * n2)
        ri_h.copy_(w.sh.data[: slabs * n2])
        mo_h = mk(slabs * n2)
        if not pinned:
            mo_h.zero_()    # touch the pages: the reference's vec![0.0; len] does the same before the FFI call
        c_h, dm_h, ct_h = mk(n2), mk(n2), mk(nb * no)
        c_h.copy_(w.c); dm_h.copy_(w.dm); ct_h.copy_(w.ct)
        d_h, j_h, k_h = mk(slabs), mk(n2), mk(n2)
        torch.cuda.synchronize()
    except Exception as exc:  # noqa: BLE001
        alloc_err = f"{type(exc).__name__}: {exc}"[:300]
    flag = torch.tensor([0.0 if alloc_err else 1.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)   # all ranks take the same branch
    if float(flag.item()) < 1.0:
        return {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "error": alloc_err or "another rank could not allocate its host buffers"}
    try:
        P = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
        jk = torch.empty(2 * n2, dtype=torch.float64, device=dev) if world > 1 else None

        def e2e_step():
            check(lib.rb_host_ri_ao2mo_jk(P(c_h), nb, P(c_h), nb, P(ri_h), P(mo_h), nb, slabs, P(dm_h), P(ct_h), no, P(d_h),
                                          P(j_h), P(k_h)), "rb_host_ri_ao2mo_jk")
            if world > 1:  # complete J and K across ranks (host results -> NVLink all-reduce -> host)
                jk[:n2].copy_(j_h, non_blocking=True); jk[n2:].copy_(k_h, non_blocking=True)
                env.ctx.allreduce_sum(jk)
                j_h.copy_(jk[:n2]); k_h.copy_(jk[n2:])
                torch.cuda.synchronize()

        e2e_step()
        # parity spot check against the device-resident results of the same inputs (before J/K get all-reduced again)
        ok = None
        if slabs == nx and w.mo is not None:
            idx = torch.arange(0, slabs * n2, max(1, slabs * n2 // 65536))
            ok = bool(torch.allclose(mo_h[idx], w.mo[idx.to(dev)].cpu(), rtol=1e-12, atol=1e-14))
        env.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            e2e_step()
        env.barrier()
        dt = env.max_over_ranks((time.perf_counter() - t0) / steps)
        flop = total_flop * slabs / nx
        out = {"value": flop / dt / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": (slabs * n2 + 2 * n2 + nb * no) * 8, "d2h_bytes_per_step": (slabs * n2 + 2 * n2 + slabs) * 8,
               "ms_per_step": dt * 1e3,
               "api": "rb_host_ri_ao2mo_jk (host-pointer C ABI; %s host buffers; 3-stream H2D|compute|D2H pipeline)"
                      % ("pinned" if pinned else "pageable"),
               "steps": steps, "matches_device_path": ok}
        if note:
            out["note"] = note
        return out
    except Exception as exc:  # noqa: BLE001
        return {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "error": f"{type(exc).__name__}: {exc}"[:300]}


def e2e_variants_leg(env, lib, check, w, steps=3):
    """Two more end-to-end forms of the same step at N = 1 (VERDICT r01 'cut e2e bytes'), both through the C ABI with pinned
    host buffers and host<->device copies inside the timed region:
      occ_vir      : rb_host_ri_ao2mo_jk with C_left = C_occ, C_right = C_vir (north-star form): ri3ao still streams up once,
                     but the ri3mo that comes back is nocc*nvir/nb^2 of the square one;
      scf_resident : ri3ao stays in HBM between SCF iterations (what REST's loop needs): per iteration D and C~ go up
                     (rb_memcpy_h2d), d_P / J / K run on the resident shard, J and K come back (rb_memcpy_d2h)."""
    import ctypes as C
    torch, dev = env.torch, env.dev
    nb, nx, no = w.nb, w.nx, w.no
    n2, nv = nb * nb, nb - no
    out = {}
    try:
        P = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
        mk = lambda n: torch.empty(n, dtype=torch.float64, pin_memory=True)  # noqa: E731
        ri_h = mk(nx * n2); ri_h.copy_(w.sh.data[: nx * n2])
        c_h, dm_h, ct_h = mk(n2), mk(n2), mk(nb * no)
        c_h.copy_(w.c); dm_h.copy_(w.dm); ct_h.copy_(w.ct)
        ov_h, d_h, j_h, k_h = mk(nx * no * nv), mk(nx), mk(n2), mk(n2)
        torch.cuda.synchronize()
        cocc = C.c_void_p(c_h.data_ptr()); cvir = C.c_void_p(c_h.data_ptr() + nb * no * 8)

        def ov_step():
            check(lib.rb_host_ri_ao2mo_jk(cocc, no, cvir, nv, P(ri_h), P(ov_h), nb, nx, P(dm_h), P(ct_h), no, P(d_h), P(j_h),
                                          P(k_h)), "rb_host_ri_ao2mo_jk(occ-vir)")
        ov_step()
        w.step(False)
        torch.cuda.synchronize()
        idx = torch.arange(0, nx * no * nv, max(1, nx * no * nv // 65536))
        ok = bool(torch.allclose(ov_h[idx], w.ov[idx.to(dev)].cpu(), rtol=1e-12, atol=1e-14)) and \
            bool(torch.allclose(k_h, w.k.cpu(), rtol=1e-11, atol=1e-12))
        t0 = time.perf_counter()
        for _ in range(steps):
            ov_step()
        dt = (time.perf_counter() - t0) / steps
        flop = sum(flops(nb, nx, no, False).values())
        out["occ_vir"] = {"ms_per_step": dt * 1e3, "value": flop / dt / 1e9, "unit": UNIT,
                          "h2d_bytes_per_step": (nx * n2 + 2 * n2 + nb * no) * 8,
                          "d2h_bytes_per_step": (nx * no * nv + 2 * n2 + nx) * 8, "matches_device_path": ok,
                          "api": "rb_host_ri_ao2mo_jk(C_occ, C_vir) (host-pointer C ABI, pinned buffers)"}
        del ov_h
        # square C, but only the a <= b pairs of ri3mo travel back (-50 % D2H)
        try:
            npair = nb * (nb + 1) // 2
            up_h = mk(nx * npair)

            def up_step():
                check(lib.rb_host_ri_ao2mo_jk_upper(P(c_h), nb, P(ri_h), P(up_h), nb, nx, P(dm_h), P(ct_h), no, P(d_h), P(j_h),
                                                    P(k_h)), "rb_host_ri_ao2mo_jk_upper")
            up_step()
            w.step(True)
            torch.cuda.synchronize()
            # pair (a, b) of slab column P  <->  w.mo[P + nx * (a + nb * b)]
            bs = torch.tensor([0, 1, nb // 3, nb // 2, nb - 1]); as_ = torch.tensor([0, 0, nb // 5, nb // 2, nb - 2])
            ok_up = True
            for a_, b_ in zip(as_.tolist(), bs.tolist()):
                a_ = min(a_, b_)
                got = up_h[nx * (b_ * (b_ + 1) // 2 + a_): nx * (b_ * (b_ + 1) // 2 + a_ + 1)]
                ref = w.mo[nx * (a_ + nb * b_): nx * (a_ + nb * b_ + 1)].cpu()
                ok_up = ok_up and bool(torch.equal(got, ref))
            ok_up = ok_up and bool(torch.allclose(k_h, w.k.cpu(), rtol=1e-11, atol=1e-12))
            t0 = time.perf_counter()
            for _ in range(steps):
                up_step()
            dt = (time.perf_counter() - t0) / steps
            flop = sum(flops(nb, nx, no, True).values())
            out["upper_packed"] = {"ms_per_step": dt * 1e3, "value": flop / dt / 1e9, "unit": UNIT,
                                   "h2d_bytes_per_step": (nx * n2 + 2 * n2 + nb * no) * 8,
                                   "d2h_bytes_per_step": (nx * npair + 2 * n2 + nx) * 8, "matches_device_path": ok_up,
                                   "api": "rb_host_ri_ao2mo_jk_upper (square C; ri3mo[P, a<=b] only: the slabs are symmetric)",
                                   "note": "value counts the flop of the square step (the device still forms the full ri3mo chunk)"}
            # ... and only mu <= nu of every (symmetric) slab travels up: about half the bytes in both directions
            try:
                def sy_step():
                    check(lib.rb_host_ri_ao2mo_jk_symm(P(c_h), nb, P(ri_h), P(up_h), nb, nx, P(dm_h), P(ct_h), no, P(d_h), P(j_h),
                                                       P(k_h)), "rb_host_ri_ao2mo_jk_symm")
                ref_up = up_h.clone(); ref_k = k_h.clone(); ref_j = j_h.clone()
                up_h.zero_()
                sy_step()
                ok_sy = bool(torch.equal(up_h, ref_up)) and bool(torch.equal(k_h, ref_k)) and bool(torch.equal(j_h, ref_j))
                del ref_up
                t0 = time.perf_counter()
                for _ in range(steps):
                    sy_step()
                dt = (time.perf_counter() - t0) / steps
                up_bytes = sum((min(c0 + 32, nb)) * min(32, nb - c0) for c0 in range(0, nb, 32)) * nx * 8
                out["symmetric_in_upper_out"] = {"ms_per_step": dt * 1e3, "value": flop / dt / 1e9, "unit": UNIT,
                                                 "h2d_bytes_per_step": up_bytes + (2 * n2 + nb * no) * 8,
                                                 "d2h_bytes_per_step": (nx * npair + 2 * n2 + nx) * 8,
                                                 "bit_identical_to_upper_packed": ok_sy,
                                                 "api": "rb_host_ri_ao2mo_jk_symm (caller guarantees symmetric slabs: mu <= nu up, a <= b down)",
                                                 "note": "value counts the flop of the square step"}
            except Exception as exc:  # noqa: BLE001
                out["symmetric_in_upper_out"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
            del up_h
        except Exception as exc:  # noqa: BLE001
            out["upper_packed"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        del ri_h
        # resident shard: only D, C~ up and J, K down per iteration
        ctx = env.ctx

        def scf_step():
            check(lib.rb_memcpy_h2d(ctx.h, P(w.dm), P(dm_h), C.c_int64(n2 * 8)), "rb_memcpy_h2d")
            check(lib.rb_memcpy_h2d(ctx.h, P(w.ct), P(ct_h), C.c_int64(nb * no * 8)), "rb_memcpy_h2d")
            w.sh.dp(w.dm, out=w.d)
            w.sh.j(w.d, out=w.j, reduce=True)
            w.sh.k(w.ct, no, out=w.k, reduce=True)
            check(lib.rb_memcpy_d2h(ctx.h, P(j_h), P(w.j), C.c_int64(n2 * 8)), "rb_memcpy_d2h")
            check(lib.rb_memcpy_d2h(ctx.h, P(k_h), P(w.k), C.c_int64(n2 * 8)), "rb_memcpy_d2h")
            ctx.sync()
        scf_step()
        t0 = time.perf_counter()
        for _ in range(10):
            scf_step()
        dt = (time.perf_counter() - t0) / 10
        f = flops(nb, nx, no, True)
        flop = f["dp"] + f["j"] + f["k"]
        # the same iteration with d_P and J from ONE read of the shard (rb_ri_dp_j: persistent cooperative kernel)
        def scf_step_1pass():
            check(lib.rb_memcpy_h2d(ctx.h, P(w.dm), P(dm_h), C.c_int64(n2 * 8)), "rb_memcpy_h2d")
            check(lib.rb_memcpy_h2d(ctx.h, P(w.ct), P(ct_h), C.c_int64(nb * no * 8)), "rb_memcpy_h2d")
            w.sh.dp_j(w.dm, out_d=w.d, out_j=w.j, reduce=True)
            w.sh.k(w.ct, no, out=w.k, reduce=True)
            check(lib.rb_memcpy_d2h(ctx.h, P(j_h), P(w.j), C.c_int64(n2 * 8)), "rb_memcpy_d2h")
            check(lib.rb_memcpy_d2h(ctx.h, P(k_h), P(w.k), C.c_int64(n2 * 8)), "rb_memcpy_d2h")
            ctx.sync()
        one_pass = None
        try:
            j_two = j_h.clone()
            scf_step_1pass()
            rel = float((j_h - j_two).abs().max() / j_two.abs().max())
            t1 = time.perf_counter()
            for _ in range(10):
                scf_step_1pass()
            dt1 = (time.perf_counter() - t1) / 10
            ev0, ev1, ev2 = env.ev(), env.ev(), env.ev()
            ev0.record()
            for _ in range(5):
                w.sh.dp(w.dm, out=w.d); w.sh.j(w.d, out=w.j, reduce=False)
            ev1.record()
            for _ in range(5):
                w.sh.dp_j(w.dm, out_d=w.d, out_j=w.j, reduce=False)
            ev2.record(); torch.cuda.synchronize()
            one_pass = {"ms_per_step": dt1 * 1e3, "value": flop / dt1 / 1e9, "j_rel_diff_vs_two_passes": rel,
                        "dp_plus_j_two_passes_ms": ev0.elapsed_time(ev1) / 5, "dp_j_single_pass_ms": ev1.elapsed_time(ev2) / 5,
                        "api": "rb_ri_dp_j (d_P and J from one read of the shard) + rb_ri_k"}
        except Exception as exc:  # noqa: BLE001
            one_pass = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        out["scf_resident"] = {"ms_per_step": dt * 1e3, "value": flop / dt / 1e9, "unit": UNIT, "single_pass_dp_j": one_pass,
                               "h2d_bytes_per_step": (n2 + nb * no) * 8, "d2h_bytes_per_step": 2 * n2 * 8,
                               "api": "rb_memcpy_h2d(D, C~) + rb_ri_dp / rb_ri_j / rb_ri_k on the resident shard + rb_memcpy_d2h(J, K)",
                               "note": "d_P + J + K only (the per-iteration work of an SCF loop); ri3ao uploaded once, outside"}
    except Exception as exc:  # noqa: BLE001
        out["error"] = f"{type(exc).__name__}: {exc}"[:300]
    return out


def crossover_leg(lib, check, threads=8, budget_s=1.0):
    """Drop-in crossover: REST calls the BLAS wrappers PER SLAB from rayon worker threads (the reference's
    par_iter_auxbas pattern, src/ri.rs:180-218).  For slab sizes nb = 100 / 264 / 600 this times, from `threads` concurrent
    caller threads, the host-pointer entry points (upload -> kernel -> download under the library's per-device mutex)
    against the same call on OpenBLAS with one BLAS thread per caller (the oracle is the CPU baseline here), and next to
    them the BATCHED entry point that does the same per-slab work for a whole block of slabs in one call (rb_host_ri_k:
    dgemm + dsyrk per slab).  Microseconds per call, wall clock / calls."""
    import ctypes as C
    import numpy as np
    from oracle.api import Oracle
    from rest_tensors_b200 import tensors as rt
    o = Oracle()
    if not o.load_openblas(threads=1):
        return {"error": "no OpenBLAS for the CPU side"}
    out = {"caller_threads": threads, "unit": "us per call (wall / calls, all caller threads running)",
           "cpu": "OpenBLAS, 1 BLAS thread per caller thread; " + o.blas_config()[:60], "rows": []}

    def timed(fn_per_thread, budget):
        """every thread loops fn until the budget is spent; returns us per call"""
        counts = [0] * threads
        stop = time.perf_counter() + budget

        def work(t):
            fn = fn_per_thread(t)
            fn()
            n = 0
            while time.perf_counter() < stop:
                fn(); n += 1
            counts[t] = n
        ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
        t0 = time.perf_counter()
        for th in ths: th.start()
        for th in ths: th.join()
        dt = time.perf_counter() - t0
        return dt / max(1, sum(counts)) * 1e6

    for nb, no in [(100, 20), (264, 21), (600, 60)]:
        n2 = nb * nb
        bufs = []
        for t in range(threads):
            a = o.fill_linear(n2, 100 + t); ct = o.fill_linear(nb * no, 200 + t, scale=nb ** -0.5)
            bufs.append({"a": a, "ct": ct, "y": np.zeros(nb * no), "k": np.zeros(n2), "x": o.fill_linear(nb, 300 + t),
                         "v": np.zeros(nb), "b": np.zeros(n2)})
        P = lambda v: C.c_void_p(v.ctypes.data)  # noqa: E731
        ops = {
            "dgemm  Y = A_P C~ [nb,nb]x[nb,no]": (
                lambda t: (lambda b=bufs[t]: check(lib.rb_host_dgemm(b"N", b"N", nb, no, nb, 1.0, P(b["a"]), nb, P(b["ct"]), nb, 0.0,
                                                                     P(b["y"]), nb), "rb_host_dgemm")),
                lambda t: (lambda b=bufs[t]: o.dgemm("N", "N", nb, no, nb, 1.0, b["a"], nb, b["ct"], nb, 0.0, b["y"], nb))),
            "dsyrk  K += Y Y^T (n=nb, k=no)": (
                lambda t: (lambda b=bufs[t]: check(lib.rb_host_dsyrk(b"U", b"N", nb, no, 1.0, P(b["y"]), nb, 1.0, P(b["k"]), nb),
                                                   "rb_host_dsyrk")),
                lambda t: (lambda b=bufs[t]: o.dsyrk("U", "N", nb, no, 1.0, b["y"], nb, 1.0, b["k"], nb))),
            "dgemv  v = A_P x [nb,nb]": (
                lambda t: (lambda b=bufs[t]: check(lib.rb_host_dgemv(b"N", nb, nb, 1.0, P(b["a"]), nb, P(b["x"]), 1, 0.0, P(b["v"]), 1),
                                                   "rb_host_dgemv")),
                lambda t: (lambda b=bufs[t]: o.dgemv("N", nb, nb, 1.0, b["a"], nb, b["x"], 1, 0.0, b["v"], 1))),
            "copy_rr one slab [nb,nb,1]": (
                lambda t: (lambda b=bufs[t]: rt.ri_copy_from_ri(b["a"], [nb, nb, 1], (0, nb), (0, nb), (0, 1), b["b"], [nb, nb, 1],
                                                                (0, nb), (0, nb), (0, 1))),
                lambda t: (lambda b=bufs[t]: o.copy_rr(nb, nb, 1, b["a"], nb, nb, 1, 0, 0, 0, b["b"], nb, nb, 1, 0, 0, 0))),
        }
        for name, (gpu, cpu) in ops.items():
            g = timed(gpu, budget_s * 0.25)
            c = timed(cpu, budget_s * 0.25)
            out["rows"].append({"nb": nb, "no": no, "op": name, "gpu_us": round(g, 1), "cpu_us": round(c, 1),
                                "gpu_over_cpu": round(g / c, 2)})
        # the batched remedy: K over a block of slabs in ONE call vs the per-slab dgemm + dsyrk on the CPU
        slabs = max(8, min(256, (64 << 20) // (n2 * 8)))
        ri = o.fill_ri3ao_symm(nb, 0, slabs)
        kk = np.zeros(n2)
        check(lib.rb_host_ri_k(P(ri), P(bufs[0]["ct"]), no, P(kk), nb, slabs), "rb_host_ri_k")
        t0 = time.perf_counter(); reps = 0
        while time.perf_counter() - t0 < budget_s * 0.25:
            check(lib.rb_host_ri_k(P(ri), P(bufs[0]["ct"]), no, P(kk), nb, slabs), "rb_host_ri_k"); reps += 1
        g = (time.perf_counter() - t0) / reps / slabs * 1e6
        o.set_threads(threads)
        t0 = time.perf_counter(); reps = 0
        while time.perf_counter() - t0 < budget_s * 0.25:
            o.ri_k(ri, bufs[0]["ct"], nb, no, slabs); reps += 1
        c = (time.perf_counter() - t0) / reps / slabs * 1e6
        o.set_threads(1)
        out["rows"].append({"nb": nb, "no": no, "op": f"BATCHED rb_host_ri_k over {slabs} slabs, per slab (dgemm + dsyrk); CPU: "
                            f"same loop, {threads} BLAS threads", "gpu_us": round(g, 1), "cpu_us": round(c, 1),
                            "gpu_over_cpu": round(g / c, 2)})
    check(lib.rb_host_trim(), "rb_host_trim")
    return out


def graph_replay_leg(env, w, flop, flush):
    """The same device-resident step recorded once into a CUDA graph (rb_graph_begin / rb_graph_end) and replayed with one launch
    per step, and the per-iteration part of an SCF loop (d_P + J + K) both ways: for the small configurations the launches are the
    cost, not the kernels.  Runs on a side stream (the default stream cannot be captured) and re-binds the context afterwards."""
    torch, ctx = env.torch, env.ctx
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    res = {}

    def scf_iter():
        w.sh.dp(w.dm, out=w.d); w.sh.j(w.d, out=w.j, reduce=True); w.sh.k(w.ct, w.no, out=w.k, reduce=True)

    def timed(fn, reps=10):
        ts = []
        for _ in range(reps):
            flush()
            e0, e1 = env.ev(), env.ev()
            e0.record(); fn(); e1.record()
            side.synchronize()
            ts.append(e0.elapsed_time(e1))
        return sorted(ts)[len(ts) // 2]

    with torch.cuda.stream(side):
        ctx.bind_stream()
        try:
            w.step(True)
            want = [t.clone() for t in (w.mo, w.d, w.j, w.k)]
            with ctx.record() as rec_step:
                w.step(True)
            with ctx.record() as rec_iter:
                scf_iter()
            for t in (w.mo, w.d, w.j, w.k):
                t.zero_()
            for _ in range(3):
                rec_step.graph.launch()
            side.synchronize()
            same = all(bool(torch.equal(a, b)) for a, b in zip((w.mo, w.d, w.j, w.k), want))
            ms_direct = timed(lambda: w.step(True))
            ms_graph = timed(rec_step.graph.launch)
            it_direct = timed(scf_iter)
            it_graph = timed(rec_iter.graph.launch)
            res = {"ms": ms_graph, "gflops": flop / (ms_graph * 1e-3) / 1e9, "ms_call_by_call_same_stream": ms_direct,
                   "kernels_per_replay": rec_step.graph.kernels, "bit_identical_to_call_by_call": same,
                   "scf_iteration_dp_j_k": {"ms_call_by_call": it_direct, "ms_graph": it_graph,
                                            "kernels_per_replay": rec_iter.graph.kernels},
                   "api": "rb_graph_begin / rb_graph_end once, rb_graph_launch per step"}
            rec_step.graph.close(); rec_iter.graph.close()
        finally:
            side.synchronize()
    ctx.bind_stream()
    return res


def small_config_leg(env, lib, check, name):
    """Configs A / B on one GPU: device-resident step, e2e through the host-pointer ABI with pinned
    and with pageable buffers, and the reference CPU path on the FULL configuration, all in GFLOP/s of the same step."""
    nb, nx, no, desc = CONFIGS[name]
    torch = env.torch
    w = Workload(env, nb, nx, no)
    flop = w.total_flop(True)
    flush = l2_flusher(env) if nx * nb * nb * 8 < (256 << 20) else (lambda: None)
    for _ in range(3):
        w.step(True)
    times = []
    for _ in range(10):
        flush()
        e0, e1 = env.ev(), env.ev()
        e0.record(); w.step(True); e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = sorted(times)[len(times) // 2]
    out = {"workload": desc, "device": {"ms": ms, "gflops": flop / (ms * 1e-3) / 1e9,
                                        "l2": "flushed between iterations" if nx * nb * nb * 8 < (256 << 20) else "inputs > L2"}}
    try:
        out["device_graph"] = graph_replay_leg(env, w, flop, flush)
    except Exception as exc:  # noqa: BLE001 -- the line must survive a refused recording
        out["device_graph"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    for pinned in (True, False):
        r = _e2e_bound(env, lib, check, w, flop, 5, pinned)
        out["e2e_pinned" if pinned else "e2e_pageable"] = {"ms": r.get("ms_per_step"), "gflops": r.get("value"),
                                                            "matches_device_path": r.get("matches_device_path"),
                                                            **({"error": r["error"]} if "error" in r else {})}
    cpu = CpuPath(nb, no)
    cpu.set_sample(nx)
    cpu.step()
    cts = []
    for _ in range(3):
        t0 = time.perf_counter(); cpu.step(); cts.append(time.perf_counter() - t0)
    cdt = sorted(cts)[1]
    out["cpu_reference"] = {"ms": cdt * 1e3, "gflops": flop / cdt / 1e9, "cores": cpu.threads, "sample": "full configuration"}
    return out


def run_ours(args):
    quiet_stdout()
    import ctypes as C  # noqa: F401
    import torch
    import torch.distributed as dist
    from rest_tensors_b200 import lib
    from rest_tensors_b200._lib import check
    from rest_tensors_b200.device import Context

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs CUDA devices (no CPU fallback)"
    torch.cuda.set_device(local)
    os.environ["REST_B200_DEVICE"] = str(local)   # the host-pointer C ABI (e2e leg) runs on this rank's GPU
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{local}"))
    ctx = Context(local)
    env = Env(torch, dist, ctx, rank, world, local)
    nb, nx, no, desc = CONFIGS[args.config]
    naux = nx * world
    w = Workload(env, nb, naux, no)
    assert w.nx == nx
    n2 = nb * nb
    f = flops(nb, nx, no)

    # roofline denominators, measured live before the timed region
    dmma_peak = max(ctx.fp64_peak_probe(0, 100000)[0] for _ in range(2))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak, hbm_src = (peaks["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (of measured)") if "hbm_gbs" in peaks else \
        (6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)")

    # ---- the timed region of `value` ----
    warm = max(args.warmup, 3)
    for _ in range(warm):
        w.step(True)
    env.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.2)
    t_wall0 = time.perf_counter()
    ms_step, avg, launches = w.measure(args.steps, 0, True)
    t_wall1 = time.perf_counter()
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    total_flop = w.total_flop(True)
    value = total_flop / (ms_step * 1e-3) / 1e9
    nccl_ver = ctx.nccl_version() if world > 1 else None

    # ---- step_occ_vir: the north star's step (occ-vir ao2mo + d_P + J + K incl. all-reduces), same shards ----
    step_occ_vir = None
    if not args.no_extras:
        ov_ms, ov_seg, _ = w.measure(max(3, min(args.steps, 5)), 2, False)
        ov_flop = w.total_flop(False)
        step_occ_vir = {"ms_per_step": ov_ms, "value": ov_flop / (ov_ms * 1e-3) / 1e9, "unit": UNIT,
                        "frac_of_dmma_peak": ov_flop / (ov_ms * 1e-3) / 1e12 / (dmma_peak * world),
                        "breakdown_ms": {kk: round(v, 4) for kk, v in ov_seg.items()},
                        "breakdown_rate": rates(nb, nx, no, ov_seg, False),
                        "note": "C_left = C_occ [nb,nocc], C_right = C_vir [nb,nb-nocc]; flop = (2 no nb^2 + 2 no nb nv) nx + K + d_P + J"}

    # ---- strong scaling of config C: 1700 slabs over the N ranks (BASELINE: 'n_aux=1700, 1-8 B200') ----
    strong = None
    if world > 1 and args.config == "C" and not args.no_extras:
        ws_ = Workload(env, nb, nx, no)
        s_ms, s_seg, s_launches = ws_.measure(max(args.steps, 10), 3, True)
        ar = allreduce_ms(env, n2)
        sf = ws_.total_flop(True)
        t1 = ms_step   # this run's weak step = the same 1700 slabs on ONE GPU (plus two all-reduces)
        crit = max(s_seg, key=s_seg.get)
        strong = {"slabs_global": nx, "slabs_this_rank": ws_.nx, "ms_per_step": s_ms, "value": sf / (s_ms * 1e-3) / 1e9,
                  "unit": UNIT, "speedup_vs_1gpu_same_run": t1 / s_ms, "efficiency": t1 / s_ms / world,
                  "frac_of_dmma_peak": sf / (s_ms * 1e-3) / 1e12 / (dmma_peak * world),
                  "breakdown_ms_rank0": {kk: round(v, 4) for kk, v in s_seg.items()},
                  "breakdown_rate_rank0": rates(nb, ws_.nx, no, s_seg, True),
                  "allreduce_ms_each": ar, "allreduce_bytes": n2 * 8, "gpu_launches_per_step": s_launches / max(args.steps, 10),
                  "largest_phase": crit,
                  "note": "1-GPU reference time = this run's weak step (1700 slabs per rank); j and k phases include their all-reduce"}
        del ws_

    # ---- parity of the multi-rank path against the oracle (reduced-slab replica, same shapes) ----
    parity = {}
    if not args.no_parity:
        try:
            parity["C_shape"] = parity_leg(env, nb, no, 16)
        except Exception as exc:  # noqa: BLE001
            parity["C_shape"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    # ---- extras on the occ-vir ri3mo of this rank (the step after ao2mo) ----
    iajb = None
    if not args.no_extras:
        iajb = consumers_leg(env, w)

    # ---- e2e through the host-pointer C ABI (H2D + D2H inside the timed region) ----
    e2e_variants = None
    if args.no_e2e:
        e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "skipped": "--no-e2e"}
        e2e_pageable = None
    else:
        e2e = e2e_leg(env, lib, check, w, total_flop, max(2, min(args.steps, 5)), True, args.no_numa)
        e2e_pageable = None
        if world == 1:
            check(lib.rb_host_trim(), "rb_host_trim")
            e2e_pageable = e2e_leg(env, lib, check, w, total_flop, 3, False, args.no_numa)
            if not args.no_extras:
                e2e_variants = e2e_variants_leg(env, lib, check, w)
        check(lib.rb_host_trim(), "rb_host_trim")

    # ---- config D (north star) on 8 GPUs: nb=1800, naux=4800, nocc=180, 600 slabs per rank ----
    config_d = None
    if (world == 8 or args.config_d_leg) and args.config == "C" and not args.no_extras:
        del w
        torch.cuda.empty_cache()
        try:
            config_d = config_d_leg(env, dmma_peak, hbm_peak, args)
        except Exception as exc:  # noqa: BLE001
            config_d = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        if not args.no_parity:
            try:
                parity["D_shape"] = parity_leg(env, 1800, 180, 4)
            except Exception as exc:  # noqa: BLE001
                parity["D_shape"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    # ---- configs A and B on one GPU ----
    small = None
    if world == 1 and args.config == "C" and not args.no_extras and not args.no_e2e:
        small = {}
        for name in ("A", "B"):
            try:
                small[name] = small_config_leg(env, lib, check, name)
            except Exception as exc:  # noqa: BLE001
                small[name] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        check(lib.rb_host_trim(), "rb_host_trim")
    crossover = None
    if world == 1 and args.config == "C" and not args.no_extras and not args.no_e2e and not args.no_cpu:
        try:
            crossover = crossover_leg(lib, check)
        except Exception as exc:  # noqa: BLE001
            crossover = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    if rank != 0:
        if world > 1:
            ctx.comm_destroy()
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

    traffic = traffic_dp = traffic_src = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        traffic, traffic_dp, traffic_src = tj.get(args.config), tj.get(args.config + "_dp"), tj.get("source")
    except Exception:
        pass
    gemm_launch_ms = avg["ao2mo"] / 2.0                      # ao2mo = 2 launches of the TMA+DMMA GEMM kernel
    achieved = (f["ao2mo"] / 2.0) / (gemm_launch_ms * 1e-3) / 1e12
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": config_dict(args.config, world),
        "breakdown_ms": {kname: round(v, 4) for kname, v in avg.items()},
        "breakdown_rate": rates(nb, nx, no, avg, True),
        "roofline": {"bound": "tensor", "kernel": "rb_gemm_tma_kernel<A_K=1,B_K=1> (ao2mo GEMMs; DMMA.8x8x4 fed by TMA)",
                     "achieved": achieved, "peak": dmma_peak, "unit": "TFLOP/s", "frac": achieved / dmma_peak,
                     "peak_source": "live register-resident DMMA.8x8x4 probe on this GPU (MEASURED_PEAKS.json holds only "
                                    "bf16/HBM; nominal B200 FP64 tensor 37-40 TFLOP/s; probe SASS + ncu pipe utilisation: "
                                    "profiles/r02_fp64_probe.md)",
                     "flop_per_launch": f["ao2mo"] / 2.0, "launch_ms": gemm_launch_ms, "traffic": traffic,
                     "traffic_source": traffic_src or "profiles/roofline_traffic.json (an ncu --set full capture of this "
                                                      "command, NOT re-measured by this run)"},
        "roofline_hbm": {"bound": "hbm", "kernel": "rb_gemv_t_vec_kernel (d_P) / rb_gemv_n_kernel (J)",
                         "achieved": nx * n2 * 8 / (avg["dp"] * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": nx * n2 * 8 / (avg["dp"] * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src,
                         "j_achieved": nx * n2 * 8 / (avg["j"] * 1e-3) / 1e9, "traffic": traffic_dp},
        "e2e": e2e,
        "e2e_pageable": e2e_pageable,
        "e2e_variants": e2e_variants,
        "step_occ_vir": step_occ_vir,
        "strong_C": strong,
        "config_D": config_d,
        "parity": parity,
        "small_configs": small,
        "crossover": crossover,
        "iajb_occ_vir": iajb,
        "collectives": None if world == 1 else {"api": "rb_allreduce_sum / rb_ri_j_allreduce / rb_ri_k_allreduce (librest_b200, NCCL "
                                                       "bound at run time, compute stream)", "nccl_version": nccl_ver[0],
                                                "nccl_lib": nccl_ver[1]},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if cpu_baseline is not None:
        line["cpu_baseline"] = cpu_baseline
    emit_json_line(line)
    if world > 1:
        ctx.comm_destroy()
        dist.destroy_process_group()
    return 0


def config_d_leg(env, dmma_peak, hbm_peak, args):
    """North-star configuration: C60/cc-pVTZ-sized ri3ao (nb=1800, naux=4800, nocc=180; 124 GB) P-sharded over the ranks;
    square ao2mo + d_P + J + K with both all-reduces, device-timed like `value`; plus the occ-vir step."""
    nb, _, no, desc = CONFIGS["D"]
    naux = 4800
    w = Workload(env, nb, naux, no)
    steps = 3
    ms, seg, launches = w.measure(steps, 2, True)
    tf = w.total_flop(True)
    f = flops(nb, w.nx, no)
    achieved = f["ao2mo"] / (seg["ao2mo"] * 1e-3) / 1e12
    out = {"workload": desc, "naux_global": naux, "slabs_this_rank": w.nx, "steps": steps, "warmup": 2,
           "ms_per_step": ms, "value": tf / (ms * 1e-3) / 1e9, "unit": UNIT,
           "frac_of_dmma_peak": tf / (ms * 1e-3) / 1e12 / (dmma_peak * env.world),
           "target": ">= 0.60 of FP64 tensor peak on 8 x B200 (BASELINE.json north_star)",
           "breakdown_ms_rank0": {kk: round(v, 4) for kk, v in seg.items()},
           "breakdown_rate_rank0": rates(nb, w.nx, no, seg, True),
           "roofline": {"bound": "tensor", "kernel": "rb_gemm_tma_kernel<1,1> (ao2mo GEMMs)", "achieved": achieved,
                        "peak": dmma_peak, "unit": "TFLOP/s", "frac": achieved / dmma_peak,
                        "flop_per_launch": f["ao2mo"] / 2.0, "launch_ms": seg["ao2mo"] / 2.0},
           "roofline_hbm": {"bound": "hbm", "achieved": w.nx * nb * nb * 8 / (seg["dp"] * 1e-3) / 1e9, "peak": hbm_peak,
                            "unit": "GB/s", "frac": w.nx * nb * nb * 8 / (seg["dp"] * 1e-3) / 1e9 / hbm_peak},
           "allreduce_ms_each": allreduce_ms(env, nb * nb), "allreduce_bytes": nb * nb * 8,
           "gpu_launches_per_step": launches / steps}
    w.mo = None
    env.torch.cuda.empty_cache()
    ov_ms, ov_seg, _ = w.measure(steps, 1, False)
    ovf = w.total_flop(False)
    out["step_occ_vir"] = {"ms_per_step": ov_ms, "value": ovf / (ov_ms * 1e-3) / 1e9, "unit": UNIT,
                           "frac_of_dmma_peak": ovf / (ov_ms * 1e-3) / 1e12 / (dmma_peak * env.world),
                           "breakdown_ms_rank0": {kk: round(v, 4) for kk, v in ov_seg.items()},
                           "breakdown_rate_rank0": rates(nb, w.nx, no, ov_seg, False)}
    return out


def consumers_leg(env, w):
    """(ia|jb) blocks and the RPA-type contraction straight from this rank's occ-vir ri3mo (outside any timed region)"""
    torch, ctx, sh = env.torch, env.ctx, w.sh
    nb, nx, no = w.nb, w.nx, w.no
    nv = nb - no
    try:
        if w.ov is None:
            w.step(False)
        ov = w.ov
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
                a0, a1 = env.ev(), env.ev()
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
            a0, a1 = env.ev(), env.ev()
            a0.record(); ctx.ri_mo_pq(ov, nx, nx, ov, nx, nx, no, nv, (0, no, 0, nv), wts, 0.0, pi, nx); a1.record()
            torch.cuda.synchronize()
            if it:
                best = a0.elapsed_time(a1) if best is None else min(best, a0.elapsed_time(a1))
        res["mo_pq_weighted"] = {"ms": best, "tflops_per_gpu": float(nx) * (nx + 1) * no * nv / (best * 1e-3) / 1e12,
                                 "m": nx, "k": no * nv}
        return {"occ_block": li, "rows": m_blk, "k": nx, **res,
                "note": "rb_ri_iajb on this rank's rows of the occ-vir ri3mo (partial sum; all-reduce not timed); best of 2"}
    except Exception as exc:  # noqa: BLE001
        return {"error": f"{type(exc).__name__}: {exc}"[:200]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-numa", action="store_true", help="e2e leg: do not bind the rank to its GPU's NUMA node")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra legs (occ-vir step, strong scaling, config D, "
                    "small configs, ri3mo consumers): the ncu launch list then holds the timed step's kernels only")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-pointer e2e legs")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle parity leg")
    ap.add_argument("--config-d-leg", action="store_true", help="run the config D leg at this N too (needs naux 4800 / N "
                    "slabs of nb=1800 plus the square ri3mo per GPU: N >= 2)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
