#!/bin/bash
mkdir -p gpurun_out
for v in base occ5 occ6; do
  if [ $v = base ]; then unset REST_B200_LIB; else export REST_B200_LIB=$PWD/tools/micro/variants/lib$v.so; fi
  timeout 300 python tools/hbm_probe.py gpurun_out/hbm_$v.json > gpurun_out/hbm_$v.log 2>&1; echo "$v rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/hbm_$v.json"))
print("$v", {k: round(v) for k, v in d.items() if any(s in k for s in ("unpack_8000","unpack_4000","transpose_8000","transpose_4000","ri_transpose","unpack_1800"))})
PY
done
