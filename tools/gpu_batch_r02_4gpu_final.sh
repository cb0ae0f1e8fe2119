#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29741 bench.py --gpus 4 > gpurun_out/bench_n4_final.json 2> gpurun_out/bench_n4_final.err
echo "bench n4 rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n4_final.json').read().strip().splitlines()[-1])
print(d['value'], d['n_gpus'], 'e2e', d['e2e']['value'])
sc=d.get('strong_C') or {}; print('strong_C', sc.get('ms_per_step'), sc.get('speedup_vs_1gpu_same_run'), sc.get('efficiency'))
print('parity ok', (d.get('parity') or {}).get('C_shape', {}).get('ok'))
PY
