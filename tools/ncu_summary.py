"""Summarise ncu output for profiles/: (a) a launch list CSV (gpu__time_duration.sum per launch) -> per-kernel share table,
(b) a `--set full` .ncu-rep -> key metrics per captured launch.  Usage:
    python tools/ncu_summary.py launches <launches.csv> <out.md>
    python tools/ncu_summary.py full <report.ncu-rep> <out.md>"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
]


def launches(path, out):
    rows = list(csv.reader(open(path)))
    hdr, data = None, []
    for r in rows:
        if r and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(dict(zip(hdr, r)))
    agg = collections.OrderedDict()
    for d in data:
        if d["Metric Name"] != "gpu__time_duration.sum":
            continue
        a = agg.setdefault(d["Kernel Name"], [0, 0.0])
        a[0] += 1
        a[1] += float(d["Metric Value"].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list summary ({path}; gpu__time_duration.sum, --clock-control none; cold-cache, serialised)\n\n")
        f.write("| launches | total ms | share | kernel |\n|---:|---:|---:|---|\n")
        for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| {a[0]} | {a[1] / 1e6:.3f} | {100 * a[1] / tot:.2f}% | `{k[:110]}` |\n")
        # the step's own kernels only: the FP64 peak probe and the synthetic-input generators run outside the timed region
        step = {k: a for k, a in agg.items() if "probe" not in k and "fill" not in k and "axpy" not in k}
        stot = sum(a[1] for a in step.values())
        f.write("\nKernels of the timed step only (probe / generator launches excluded):\n\n")
        f.write("| launches | total ms | share of step | kernel |\n|---:|---:|---:|---|\n")
        for k, a in sorted(step.items(), key=lambda x: -x[1][1]):
            f.write(f"| {a[0]} | {a[1] / 1e6:.3f} | {100 * a[1] / stot:.2f}% | `{k[:110]}` |\n")


def full(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary ({path})\n")
        for r in rows[2:]:
            f.write("\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in idx:
                    f.write(f"| {k} | {r[idx[k]][:120]} | {units[idx[k]]} |\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
