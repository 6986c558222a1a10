#!/bin/bash
set -u
TAG=${1:-r2g}
mkdir -p gpurun_out
for rep in 1 2 3; do
for unit in 2048,2048 2048,4096 2048,8192; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-extra --cfg4 0 --unit $unit > gpurun_out/${TAG}_u.json 2>/dev/null
  python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_u.json")); r = d["roofline"]
print("unit $unit rep $rep:", round(d["ms_per_step"], 2), "ms; mma", round(r["kernel_ms"], 2), "e2e", round(d["e2e"]["ms_per_step"], 1), "clocks", d["clocks"]["sm_mhz"], d["clocks"]["power_w_max"], d["clocks"]["reasons"])
PY
done
done 2>&1 | tee gpurun_out/${TAG}_units_sustained.log
