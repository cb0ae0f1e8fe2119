#!/bin/bash
# stream-K validation: contraction tests (incl. repeatability + stream-K shapes), eig / iajb (they use the GEMM core), K timings with and
# without stream-K, config-E sweep, short bench
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests/test_gpu_contractions.py tests/test_gpu_iajb.py tests/test_gpu_eig.py -m gpu -q -p no:cacheprovider -x > gpurun_out/pytest_sk.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/pytest_sk.log
for sk in 0 1; do
  for cfg in "600 1700 60" "264 720 21" "1800 600 180" "100 400 20"; do
    REST_B200_STREAMK=$sk timeout -k 10 300 python tools/prof_k.py $cfg 2>&1 | tail -1 | sed "s/^/streamk=$sk /"
  done
done | tee gpurun_out/k_timings_sk.txt
for sk in 0 1; do echo "== sweep E streamk=$sk"; REST_B200_STREAMK=$sk timeout -k 10 300 python tools/sweep_e.py 2>&1 | grep '"n"' | cut -c1-330; done | tee gpurun_out/sweep_sk.txt
timeout -k 10 300 python tools/syrk_repro.py 2>&1 | tail -12
