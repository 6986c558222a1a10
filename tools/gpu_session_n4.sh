#!/bin/bash
set -u
TAG=${1:-r2m4}; N=${2:-4}
mkdir -p gpurun_out
timeout 900 python bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_default.json 2> gpurun_out/${TAG}_default.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_default.json"))
    r = d["roofline"]
    print("default:", round(d["ms_per_step"], 2), "ms", round(d["value"] / 1e3, 1), "TFLOP/s; mma", round(r["kernel_ms"], 2), "ms; parity", d["parity"], "e2e", d["e2e"]["ms_per_step"])
    print("   ", d["config"]["workload"][:330])
    for k, v in d["extra"].items():
        print("   ", k, json.dumps(v)[:700])
except Exception as e:
    print("parse failed", e)
    import subprocess; print(subprocess.run(["tail", "-15", "gpurun_out/${TAG}_default.err"], capture_output=True, text=True).stdout)
PY
