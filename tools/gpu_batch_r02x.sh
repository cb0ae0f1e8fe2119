#!/bin/bash
mkdir -p gpurun_out
for cfg in "100 400 20" "264 720 21"; do
  python tools/prof_step.py $cfg
  tag=$(echo $cfg | tr ' ' '_')
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/step_launches_$tag.csv python tools/prof_step.py $cfg > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/step_launches_$tag.csv")) if len(r)>10 and r[0].isdigit()]
# last step = last N kernels, find N by the period of names
names=[r[4] for r in rows]; durs=[float(r[-1]) for r in rows]
units=rows[0][-2] if rows else ''
# a step's kernel count: distance between the last two occurrences of the first kernel name of a step is unknown; print the last 12
for n,d in list(zip(names,durs))[-12:]:
    print(f"   {d:10.1f} {units}  {n[:90]}")
PY
done
