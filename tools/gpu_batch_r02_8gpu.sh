#!/bin/bash
# 8 x B200: multi-GPU parity tests (worlds 2, 4, 8 through the C ABI's NCCL collectives), bench at N = 8 (default flags: what
# the driver runs; includes strong_C, config_D, parity legs) and N = 4 (device-timed legs only)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | head -9 > gpurun_out/gpus8.txt
timeout -k 10 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_dist8.log 2>&1
echo "pytest dist rc=$?"; tail -4 gpurun_out/pytest_dist8.log
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29731 bench.py --gpus 8 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
echo "bench n8 rc=$?"; cut -c1-300 gpurun_out/bench_n8.json; tail -3 gpurun_out/bench_n8.err
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29741 bench.py --gpus 4 --no-e2e > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err
echo "bench n4 rc=$?"; cut -c1-300 gpurun_out/bench_n4.json; tail -3 gpurun_out/bench_n4.err
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29751 bench.py --gpus 2 --no-e2e > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench n2 rc=$?"; cut -c1-300 gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err
