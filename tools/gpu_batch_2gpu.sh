#!/bin/bash
# gpurun --gpus 2: multi-GPU parity (NCCL all-reduce of J/K/(ia|jb), peer-panel mo_pq over NVLink), peer probe, 2-GPU bench.
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout -k 10 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_dist.log 2>&1
echo "pytest dist rc=$?"; tail -15 gpurun_out/pytest_dist.log
echo skip peer probe

timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 \
    bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench2 rc=$?"; cut -c1-300 gpurun_out/bench_n2.json
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 \
    tools/mo_pq_dist_probe.py > gpurun_out/mo_pq_dist.log 2>&1
echo "mo_pq dist rc=$?"; tail -3 gpurun_out/mo_pq_dist.log
