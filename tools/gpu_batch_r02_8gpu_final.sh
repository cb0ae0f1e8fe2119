#!/bin/bash
# 8 x B200, final build of round 2: multi-GPU parity tests (worlds 2, 4, 8; incl. the recorded d_P + J + K iteration with both all-reduces and the
# single-pass d_P + J on a shard) and the default N = 8 bench line (what the driver runs: weak C, strong_C, config_D, parity)
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -p no:cacheprovider --tb=short > gpurun_out/pytest_dist8_final.log 2>&1
echo "pytest dist rc=$?"; tail -4 gpurun_out/pytest_dist8_final.log
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29731 bench.py --gpus 8 > gpurun_out/bench_n8_final.json 2> gpurun_out/bench_n8_final.err
echo "bench n8 rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n8_final.json').read().strip().splitlines()[-1])
print(d['value'], d['n_gpus'], 'e2e', d['e2e']['value'])
sc=d.get('strong_C') or {}; print('strong_C', sc.get('ms_per_step'), sc.get('speedup_vs_1gpu_same_run'), sc.get('efficiency'))
cd=d.get('config_D') or {}; print('config_D', cd.get('value'), (cd.get('roofline') or {}).get('frac'))
print('parity', d.get('parity'))
PY
tail -2 gpurun_out/bench_n8_final.err
