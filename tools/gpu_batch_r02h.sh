#!/bin/bash
mkdir -p gpurun_out
for v in C D B; do
  echo "== variant $v"
  REST_B200_LIB=$PWD/tools/micro/variants/lib$v.so REST_B200_SPLIT_ORDER=0 REST_B200_FUSED_SPLITK=0 timeout -k 10 600 python tools/syrk_repro.py 2>&1 | tail -14
done 2>&1 | tee gpurun_out/syrk_repro_variants2.txt
