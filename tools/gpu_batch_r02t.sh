#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short > gpurun_out/pytest_r02t.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/pytest_r02t.log
