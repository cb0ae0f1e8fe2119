#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python bench.py > gpurun_out/bench_r02aq.json 2> gpurun_out/bench_r02aq.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02aq.json').read().strip().splitlines()[-1])
print("ours:", d['value'], "e2e", d['e2e']['value'], d['e2e']['ms_per_step'], "frac", d['roofline']['frac'])
print(json.dumps(d['e2e_variants'].get('scf_resident'), indent=1)[:900])
PY
