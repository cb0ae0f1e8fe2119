#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/e2e_symm_probe.py > gpurun_out/e2e_symm_probe.txt 2>&1; echo "rc=$?"
grep "^==" gpurun_out/e2e_symm_probe.txt
python - <<'PY'
# last traced call of each mode: print its chunk lines
import re
txt=open('gpurun_out/e2e_symm_probe.txt').read()
blocks=txt.split("== ")
for b in blocks:
    lines=b.splitlines()
    if not lines: continue
    ch=[l for l in lines if 'chunk' in l and 'H2D' in l]
    if ch:
        print("--", lines[0][:40]); print("\n".join(ch[-7:]))
PY
