#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_layout.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
timeout -k 10 300 python tools/hbm_probe.py gpurun_out/hbm_256c.json > gpurun_out/hbm_256c.log 2>&1; echo "hbm rc=$?"; cat gpurun_out/hbm_256c.json
timeout -k 10 600 python bench.py --no-e2e --no-cpu --no-extras > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_short.json').read().strip().splitlines()[-1])
print(d['value'], d['breakdown_ms'], d['breakdown_rate'])
PY
