#!/bin/bash
# 2 x B200 sanity of the final build: multi-GPU parity tests that fit two GPUs + the default N = 2 bench line (what the driver runs)
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_dist2.log 2>&1
echo "pytest dist rc=$?"; tail -4 gpurun_out/pytest_dist2.log
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29751 bench.py --gpus 2 > gpurun_out/bench_n2_final.json 2> gpurun_out/bench_n2_final.err
echo "bench n2 rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2_final.json').read().strip().splitlines()[-1])
print(d['value'], d['n_gpus'], d['e2e']['value'], d.get('strong_C'), d.get('parity'))
PY
tail -3 gpurun_out/bench_n2_final.err
