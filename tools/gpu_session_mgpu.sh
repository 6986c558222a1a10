#!/bin/bash
# Multi-GPU session.  usage: gpurun --gpus N --timeout T -- 'bash tools/gpu_session_mgpu.sh TAG N'
set -u
TAG=${1:-r2m}; N=${2:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
echo "== mgpu tests"; timeout 900 python -m pytest tests/test_gpu_multi.py -q --timeout 800 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest.log
grep -h "mgpu_worker\|FAILS" gpurun_out/mgpu_worker_w*.log | head -20
run() { # name, extra args
  local name=$1; shift
  timeout 900 python bench.py --gpus $N --steps 5 --warmup 3 "$@" > gpurun_out/${TAG}_${name}.json 2> gpurun_out/${TAG}_${name}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_${name}.json"))
    r = d["roofline"]
    print("${name}:", round(d["ms_per_step"], 2), "ms", round(d["value"] / 1e3, 1), "TFLOP/s; mma", round(r["kernel_ms"], 2), "ms; parity", d["parity"]["mismatches"], "/", d["parity"]["checked_entries"], "gather wrong", d["parity"]["gathered_blocks_wrong"])
    print("   ", d["config"]["workload"][:330])
    for k, v in d["extra"].items():
        print("   ", k, json.dumps(v)[:600])
except Exception as e:
    print("${name}: parse failed", e)
    import subprocess; print(subprocess.run(["tail", "-15", "gpurun_out/${TAG}_${name}.err"], capture_output=True, text=True).stdout)
PY
}
run default
run whole_bcast --bcast whole --no-extra
run peer_nomc --no-multicast --no-extra
run nccl --gather nccl --bcast whole --no-extra
