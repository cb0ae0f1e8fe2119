#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_graph.py tests/test_gpu_contractions.py -m gpu -q -p no:cacheprovider --tb=short > gpurun_out/pytest_r02w.log 2>&1
echo "pytest rc=$?"; tail -40 gpurun_out/pytest_r02w.log
timeout -k 10 900 compute-sanitizer --tool initcheck --print-limit 5000 --error-exitcode 7 python tools/sanitize_target.py > gpurun_out/initcheck.log 2>&1
echo "initcheck rc=$?"; grep -c "Uninitialized" gpurun_out/initcheck.log; grep -A3 "Uninitialized" gpurun_out/initcheck.log | grep -E " at " | sed 's/+0x[0-9a-f]*//' | sort | uniq -c | sort -rn | head -30; tail -3 gpurun_out/initcheck.log
timeout -k 10 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02w.json 2> gpurun_out/bench_r02w.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02w.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['roofline']['frac'])
for k,v in (d.get('small_configs') or {}).items():
    print(k, json.dumps(v.get('device')), json.dumps(v.get('device_graph')))
PY
