#!/bin/bash
# One GPU session: tests, bench, launch list.  usage (from the repo root): gpurun --timeout 2400 -- 'bash tools/gpu_session.sh TAG'
# Outputs land in gpurun_out/ (copy what should be judged into profiles/).
set -u
TAG=${1:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
echo "== tensor-path tests"; timeout 900 python -m pytest tests/test_gpu_ozaki.py -q --timeout 300 2>&1 | tail -25 | tee gpurun_out/${TAG}_pytest_ozaki.log
echo "== all gpu tests"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -25 | tee gpurun_out/${TAG}_pytest_all.log
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -5 gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench.json"))
    r = d["roofline"]
    print("headline", round(d["ms_per_step"], 3), "ms", round(d["value"] / 1e3, 2), "TFLOP/s; mma", round(r["kernel_ms"], 2), "ms frac", round(r["frac"], 3), "whole", round(r["whole_step_frac"], 3), "e2e", round(d["e2e"]["ms_per_step"], 2), "ms; parity", d["parity"])
    for k, v in d["extra"].items():
        print(k, json.dumps(v)[:400])
except Exception as e:
    print("bench parse failed", e)
PY
echo "== launch list"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_|qb" -c 120 --csv --log-file gpurun_out/${TAG}_launches.csv python tools/ncu_qgemm.py 8192 D113 2 > /dev/null 2>&1
python profiles/summarize.py launches gpurun_out/${TAG}_launches.csv gpurun_out/${TAG}_launches.txt && tail -20 gpurun_out/${TAG}_launches.txt
