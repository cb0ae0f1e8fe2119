#!/bin/bash
mkdir -p gpurun_out
REST_B200_SPLIT_ORDER=0 REST_B200_FUSED_SPLITK=0 timeout -k 10 600 python tools/syrk_repro.py 2>&1 | tee gpurun_out/syrk_repro.txt | tail -40
