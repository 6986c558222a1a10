#!/bin/bash
# Last single-GPU session of round 2: every GPU test, smoke(), the bench line under the driver's protocol, the reference arm.
# usage: gpurun --timeout 1800 -- 'bash tools/gpu_session_final.sh TAG'
set -u
TAG=${1:-r2final}
mkdir -p gpurun_out
echo "== all gpu tests"; timeout 1200 python -m pytest tests -m gpu -q --timeout 600 2>&1 | grep -v "^    \|^$" | tail -12 | tee gpurun_out/${TAG}_pytest_all.log
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench.json"))
    r = d["roofline"]
    print("headline", round(d["ms_per_step"], 3), "ms", round(d["value"] / 1e3, 2), "TFLOP/s; mma", round(r["kernel_ms"], 2), "ms frac", round(r["frac"], 3), "whole", round(r["whole_step_frac"], 3), "e2e", round(d["e2e"]["ms_per_step"], 2), "ms; parity", d["parity"]["mismatches"], d["parity"]["checked_entries"], "clocks", d["clocks"])
    for k, v in d["extra"].items():
        print(k, json.dumps(v)[:260])
except Exception as e:
    print("bench parse failed", e)
PY
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-300
