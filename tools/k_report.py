"""profiles/r02_k_profile.md from the ncu launch lists of tools/prof_k.py (gpurun_out/r02_k_launches_{C,B,D}.csv) and the event timings."""
import collections
import csv


def load(f):
    rows = list(csv.reader(open(f)))
    hdr, data = None, []
    for r in rows:
        if r and r[0] == 'ID':
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(dict(zip(hdr, r)))
    byid = collections.OrderedDict()
    for d in data:
        byid.setdefault(d['ID'], {'k': d['Kernel Name'], 'grid': d['Grid Size']})[d['Metric Name']] = float(d['Metric Value'].replace(',', ''))
    return list(byid.values())


out = ["# Round 2 — K build per kernel (ncu launch lists, `tools/prof_k.py`, one B200, `--clock-control none`)\n",
       "`K = sum_P (A_P C~)(A_P C~)^T`: a batched `'N','N'` GEMM `Y_P = A_P C~` (packed-M tiles), ONE SYRK `'U','N'` over the stacked Y (block-list diagonal",
       "tiles; stream-K or split-K, whichever the simulated makespan favours), the fixed-order reduction of the partials, and the mirror. Times are",
       "`gpu__time_duration.sum` of the LAST of the captured repetitions; CUDA-event totals of the same calls (not under ncu) in the heading.\n"]
ev = {l.split(':')[0].split('K ')[1]: l.split(':')[1].strip() for l in open('gpurun_out/r02_k_timings.txt')}
for cfg, f, key, flop in [('C  nb=600 nx=1700 no=60', 'gpurun_out/r02_k_launches_C.csv', 'nb=600 nx=1700 no=60', (2 * 600 * 600 * 60 + 600 * 601 * 60) * 1700),
                          ('B  nb=264 nx=720 no=21', 'gpurun_out/r02_k_launches_B.csv', 'nb=264 nx=720 no=21', (2 * 264 * 264 * 21 + 264 * 265 * 21) * 720),
                          ('D shard  nb=1800 nx=600 no=180', 'gpurun_out/r02_k_launches_D.csv', 'nb=1800 nx=600 no=180', (2 * 1800 * 1800 * 180 + 1800 * 1801 * 180) * 600)]:
    L = [x for x in load(f) if 'fill' not in x['k']]
    # the last repetition = the kernels after the second-to-last symmetrize
    idx = [i for i, x in enumerate(L) if 'symmetrize' in x['k']]
    last = L[idx[-2] + 1: idx[-1] + 1]
    out.append(f"## config {cfg}   (events: {ev.get(key, '?')})\n")
    out.append("| kernel | grid | time (us) | DMMA pipe % of active | DRAM read MB | DRAM write MB | L2 hit % |\n|---|---|---:|---:|---:|---:|---:|")
    tot = 0
    for x in last:
        t = x['gpu__time_duration.sum'] / 1e3
        tot += t
        name = x['k'].replace('void <unnamed>::', '').replace('<unnamed>::', '')[:62]
        out.append(f"| `{name}` | {x['grid']} | {t:.1f} | {x.get('sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active', 0):.1f} | "
                   f"{x['dram__bytes_read.sum'] / 1e6:.0f} | {x['dram__bytes_write.sum'] / 1e6:.0f} | {x.get('lts__t_sector_hit_rate.pct', 0):.1f} |")
    out.append(f"\nSum of the kernels {tot:.0f} us -> {flop / tot / 1e6:.2f} TFLOP/s of algorithmic flop ({flop:.3e}).\n")
tail = open('profiles/r02_k_profile.md').read()
marker = "## Reading"
out.append(tail[tail.index(marker):] if marker in tail else "")
open('profiles/r02_k_profile.md', 'w').write("\n".join(out) + "\n")
