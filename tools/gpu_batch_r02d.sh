#!/bin/bash
# bisect the K additivity failure over (split order, fused reduction); TMA copy diagnostics; eig re-check
mkdir -p gpurun_out
for combo in "0 0" "1 0" "0 1" "1 1"; do
  set -- $combo
  echo "== split_order=$1 fused=$2"
  REST_B200_SPLIT_ORDER=$1 REST_B200_FUSED_SPLITK=$2 timeout -k 10 600 python -m pytest tests/test_gpu_contractions.py -m gpu -q -p no:cacheprovider -k "full_size or split or determin or triangular" 2>&1 | tail -4
done 2>&1 | tee gpurun_out/bisect.txt
timeout -k 10 600 python -m pytest tests/test_gpu_eig.py -m gpu -q -p no:cacheprovider 2>&1 | tail -4
timeout -k 10 300 python tools/tma_copy_debug.py 2>&1 | tail -30
for combo in "0 0" "1 0" "1 1"; do
  set -- $combo
  for cfg in "600 1700 60" "264 720 21" "1800 600 180" "100 400 20"; do
    REST_B200_SPLIT_ORDER=$1 REST_B200_FUSED_SPLITK=$2 timeout -k 10 300 python tools/prof_k.py $cfg 2>&1 | tail -1 | sed "s/^/split_order=$1 fused=$2 /"
  done
done | tee gpurun_out/k_timings.txt
