#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short > gpurun_out/pytest_final.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_final.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
