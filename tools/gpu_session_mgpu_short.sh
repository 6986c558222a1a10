#!/bin/bash
# Short multi-GPU check: the world-N test and the default bench line as the driver launches it.
# usage: gpurun --gpus N --timeout T -- 'bash tools/gpu_session_mgpu_short.sh TAG N'
set -u
TAG=${1:-r2m}; N=${2:-2}
mkdir -p gpurun_out
echo "== mgpu tests"; timeout 900 python -m pytest tests/test_gpu_multi.py -q --timeout 800 2>&1 | tail -6 | tee gpurun_out/${TAG}_pytest.log
grep -h "mgpu_worker\|FAILS" gpurun_out/mgpu_worker_w*.log | head -8
timeout 870 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("bench:", round(d["ms_per_step"], 2), "ms", round(d["value"] / 1e3, 1), "TFLOP/s; mma", round(r["kernel_ms"], 2), "ms; parity", d["parity"]["mismatches"], "/", d["parity"]["checked_entries"], "gather wrong", d["parity"]["gathered_blocks_wrong"], "clocks", d["clocks"])
    for k, v in d["extra"].items():
        print("   ", k, json.dumps(v)[:500])
except Exception as e:
    print("parse failed", e)
    import subprocess; print(subprocess.run(["tail", "-15", "gpurun_out/${TAG}_bench.err"], capture_output=True, text=True).stdout)
PY
