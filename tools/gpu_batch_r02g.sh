#!/bin/bash
mkdir -p gpurun_out
for v in A B; do
  echo "== variant $v"
  REST_B200_LIB=$PWD/tools/micro/variants/lib$v.so REST_B200_SPLIT_ORDER=0 REST_B200_FUSED_SPLITK=0 timeout -k 10 600 python tools/syrk_repro.py 2>&1 | tail -14
done 2>&1 | tee gpurun_out/syrk_repro_variants.txt
echo "== current under compute-sanitizer (racecheck, then initcheck) on a small failing shape"
cat > /tmp/small.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
from rest_tensors_b200.device import Context
ctx = Context(0)
n, k = 776, 20000
a = ctx.empty(n * k); ctx.fill_linear(a, n * k, 7, 0, 1.0)
outs = []
for rep in range(3):
    c = ctx.empty(n * n); c.fill_(0.0)
    ctx.dsyrk("U", "N", n, k, 1.0, a, n, 0.0, c, n)
    outs.append(c.clone())
torch.cuda.synchronize()
print("mismatch", [int((o != outs[0]).sum()) for o in outs[1:]])
PY
REST_B200_SPLIT_ORDER=0 REST_B200_FUSED_SPLITK=0 timeout -k 10 900 compute-sanitizer --tool racecheck --racecheck-report all python /tmp/small.py 2>&1 | grep -v "^=========     at\|^=========         in\|Host Frame" | head -60 | tee gpurun_out/racecheck.txt
REST_B200_SPLIT_ORDER=0 REST_B200_FUSED_SPLITK=0 timeout -k 10 900 compute-sanitizer --tool initcheck python /tmp/small.py 2>&1 | head -40 | tee gpurun_out/initcheck.txt
