#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_contractions.py -m gpu -q -p no:cacheprovider --tb=short -k "fused_streaming" > gpurun_out/pytest_r02ad.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/pytest_r02ad.log
timeout -k 10 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02ad.json 2> gpurun_out/bench_r02ad.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02ad.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'])
for k,v in (d.get('e2e_variants') or {}).items():
    print(k, {kk: vv for kk, vv in v.items() if kk not in ('api','note','unit')} if isinstance(v, dict) else v)
PY
