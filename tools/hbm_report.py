"""profiles/r02_hbm_kernels.md from the JSON tools/hbm_probe.py writes.  Usage: python tools/hbm_report.py <probe.json> <out.md>"""
import json
import sys

d = json.load(open(sys.argv[1]))
peak = 6557.8
try:
    peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    pass
out = ["# Round 2 — HBM-bound kernels (tools/hbm_probe.py, one B200, CUDA events, inputs > L2)\n",
       f"Denominator: MEASURED_PEAKS.json `hbm_gbs` = {peak} GB/s (the driver's torch `copy_` = cudaMemcpyAsync D2D). `tools/micro/copy_bench.cu` shows that this is",
       "close to the ceiling of the part for SM kernels as well: 32-byte grid-stride kernels reach 6.6-6.7 TB/s for a copy, 6.5 for a 1 read : 2 write mix, 7.0 read-only (`r02_copy_microbench.md`; its first version reported 7.2-8.2 TB/s from loops that dropped their tail),",
       "so fractions above 1.0 are possible.\n",
       "Two timings per kernel: **single call** = one launch inside one CUDA-event pair, best of 7 (includes the ~5 us an event pair adds around one launch:",
       "3 % at n = 8000, 12 % at n = 4000); **stream** = the call issued back to back over >= 1 GB of distinct buffer sets inside ONE event pair (the sustained",
       "rate of the kernel including its own launch gap).\n",
       "| kernel | algorithmic bytes | n | single call GB/s | frac | stream GB/s | frac | round 1 (single) |\n|---|---|---:|---:|---:|---:|---:|---:|"]
r01 = {'unpack_8000': 5762, 'unpack_4000': 5098, 'pack_8000': 5872, 'pack_4000': 5306, 'transpose_8000': 5876, 'copy_mm_8000': 5824, 'axpy_8000': 6472}
for kind, alg in [("unpack", "(np+n^2)*8"), ("pack", "2*np*8"), ("transpose", "2*n^2*8"), ("copy_mm", "2*n^2*8"), ("axpy", "3*n^2*8")]:
    for n in (8000, 4000, 1800):
        s = d.get(f"{kind}_{n}_gbs"); st = d.get(f"{kind}_{n}_stream_gbs")
        out.append(f"| {kind} | {alg} | {n} | {s:.0f} | {s / peak:.3f} | " + (f"{st:.0f} | {st / peak:.3f}" if st else "- | -") + f" | {r01.get(f'{kind}_{n}', '-')} |")
out.append("")
out.append("| kernel | algorithmic bytes | GB/s | frac | round 1 |\n|---|---|---:|---:|---:|")
for name, key, alg, old in [("flat copy, 1 GiB in + 1 GiB out (the micro-benchmark's footprint)", "copy_flat_1GiB_gbs", "2n*8", "-"),
                            ("matrix transpose 11584^2 (1 GiB)", "transpose_1GiB_gbs", "2n^2*8", "-"),
                            ("RIFull transpose_jik 600x600x400", "ri_transpose_jik_gbs", "2*IJK*8", 5850), ("transpose_jki", "ri_transpose_jki_gbs", "2*IJK*8", 5864),
                            ("transpose_kji", "ri_transpose_kji_gbs", "2*IJK*8", 5481), ("transpose_ikj", "ri_transpose_ikj_gbs", "2*IJK*8", 5309),
                            ("rifull_to_matfull_symm 600x600x400", "ri_pack_symm_gbs", "2*np*K*8", 6300),
                            ("copy_rr 500x520x300 box at (50,40,30) -> (20,10,60) (corners not 32-byte aligned: 16-byte kernel)", "copy_rr_box_gbs", "2*box*8", 5579),
                            ("copy probe (16-byte, 4 loads in flight)", "hbm_copy_probe_gbs", "2n*8", 6115), ("cudaMemcpy D2D via torch copy_", "torch_copy_gbs", "2n*8", 6521)]:
    if key in d:
        out.append(f"| {name} | {alg} | {d[key]:.0f} | {d[key] / peak:.3f} | {old} |")
tail = open(sys.argv[2]).read() if len(sys.argv) > 2 else ""
marker = "\nd_P / J (one pass"
keep = tail[tail.index(marker):] if marker in tail else ""
open(sys.argv[2], "w").write("\n".join(out) + "\n" + keep)
