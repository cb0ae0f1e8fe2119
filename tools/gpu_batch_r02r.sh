#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python tools/gemm_floor_probe.py 2>&1 | tee gpurun_out/gemm_floor.txt
timeout -k 10 300 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,smsp__cycles_active.avg,smsp__inst_executed.sum,sm__warps_active.avg.per_cycle_active --clock-control none -k regex:rb_gemm_tma -c 80 --csv --log-file gpurun_out/gemm_floor_ncu.csv python tools/gemm_floor_probe.py > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/gemm_floor_ncu.csv')))
hdr=None; data=[]
for r in rows:
    if r and r[0]=='ID': hdr=r; continue
    if hdr and len(r)==len(hdr): data.append(dict(zip(hdr,r)))
byid=collections.OrderedDict()
for d in data: byid.setdefault(d['ID'],{'grid':d['Grid Size']})[d['Metric Name']]=d['Metric Value']
seen=set()
for i,v in byid.items():
    key=(v['grid'],v.get('smsp__inst_executed.sum'))
    if key in seen: continue
    seen.add(key); print(i,v)
PY
