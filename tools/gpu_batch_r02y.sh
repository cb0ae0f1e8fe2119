#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_contractions.py -m gpu -q -p no:cacheprovider --tb=short -k "fused_streaming" > gpurun_out/pytest_r02y.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/pytest_r02y.log
timeout -k 10 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02y.json 2> gpurun_out/bench_r02y.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02y.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'])
print(json.dumps(d.get('e2e_variants'), indent=1))
PY
