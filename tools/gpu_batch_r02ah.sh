#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short > gpurun_out/pytest_r02ah.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/pytest_r02ah.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout -k 10 900 python bench.py --impl reference > gpurun_out/bench_ref_r02ah.json 2> gpurun_out/bench_ref_r02ah.err; echo "ref rc=$?"
timeout -k 10 900 python bench.py > gpurun_out/bench_r02ah.json 2> gpurun_out/bench_r02ah.err
echo "bench rc=$?"; python - <<'PY'
import json
r=json.loads(open('gpurun_out/bench_ref_r02ah.json').read().strip().splitlines()[-1])
print("reference arm:", r.get('value'), r.get('unit'), r.get('cpu_baseline',{}).get('cores'))
d=json.loads(open('gpurun_out/bench_r02ah.json').read().strip().splitlines()[-1])
print("ours:", d['value'], "e2e", d['e2e']['value'], d['e2e']['ms_per_step'], "frac", d['roofline']['frac'], "launches", d['gpu_launches'], "clocks", d.get('clocks'))
print("keys:", sorted(d.keys()))
PY
