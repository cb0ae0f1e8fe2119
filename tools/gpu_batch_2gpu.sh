#!/bin/bash
# gpurun --gpus 2: multi-GPU parity (NCCL all-reduce of J/K/(ia|jb), peer pipelines over NVLink) and the dist probes.
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_dist.log 2>&1
echo "pytest dist rc=$?"; tail -15 gpurun_out/pytest_dist.log
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1 --master-port 29712 \
    tools/mo_pq_dist_probe.py > gpurun_out/mo_pq_dist.log 2>&1
echo "mo_pq dist rc=$?"; tail -2 gpurun_out/mo_pq_dist.log
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1 --master-port 29713 \
    tools/special_dgemm_dist_probe.py > gpurun_out/special_dgemm_dist.log 2>&1
echo "special_dgemm dist rc=$?"; tail -2 gpurun_out/special_dgemm_dist.log
