#!/bin/bash
mkdir -p gpurun_out
export PROBE_SCHED='[["",256],["",96],["",128],["",144],["",160],["",176],["",192],["",208],["",224],["48",160],["64",192],["48,112",192],["",256]]'
timeout 900 python tools/e2e_head_probe.py > gpurun_out/e2e_pc_probe.txt 2>&1; echo "rc=$?"; cat gpurun_out/e2e_pc_probe.txt | tail -16
