#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short > gpurun_out/pytest_r02ap.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/pytest_r02ap.log
timeout -k 10 200 python tools/prof_dpj.py 2>&1 | tee gpurun_out/prof_dpj_final.txt | tail -6
REST_B200_DPJ_FUSED= timeout -k 10 900 python bench.py > gpurun_out/bench_r02ap.json 2> gpurun_out/bench_r02ap.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02ap.json').read().strip().splitlines()[-1])
print("ours:", d['value'], "e2e", d['e2e']['value'], d['e2e']['ms_per_step'], "frac", d['roofline']['frac'])
print(json.dumps(d['e2e_variants'].get('scf_resident'), indent=1)[:1500])
for k,v in (d.get('small_configs') or {}).items(): print(k, v.get('device'), v.get('e2e_pinned'))
PY
