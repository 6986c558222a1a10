#!/bin/bash
set -u
TAG=${1:-r2n2}; N=${2:-2}
mkdir -p gpurun_out
echo "== mgpu tests"; timeout 600 python -m pytest tests/test_gpu_multi.py -q --timeout 500 2>&1 | tail -12 | tee gpurun_out/${TAG}_pytest.log
grep -h "mgpu_worker\|FAILS\|Error\|error" gpurun_out/mgpu_worker_w*.log | head -20
run() { local name=$1; shift
  timeout 600 python bench.py --gpus $N --steps 10 --warmup 3 "$@" > gpurun_out/${TAG}_${name}.json 2> gpurun_out/${TAG}_${name}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_${name}.json"))
    r = d["roofline"]
    print("${name}:", round(d["ms_per_step"], 2), "ms", round(d["value"] / 1e3, 1), "TFLOP/s; mma", round(r["kernel_ms"], 2), "ms; parity", d["parity"]["mismatches"], "/", d["parity"]["checked_entries"], "gather wrong", d["parity"]["gathered_blocks_wrong"])
    print("   ", d["config"]["workload"][-330:])
    c = d["extra"].get("cfg4_strong")
    if c: print("    cfg4:", c.get("ms_per_step"), c.get("tensor_kernel_ms"), c.get("parity"), c.get("error"), c.get("workload", "")[-200:])
except Exception as e:
    print("${name}: parse failed", e)
    import subprocess; print(subprocess.run(["tail", "-25", "gpurun_out/${TAG}_${name}.err"], capture_output=True, text=True).stdout)
PY
}
run shared --no-extra
run noshare --no-extra --no-share-planes
run shared_cfg4
