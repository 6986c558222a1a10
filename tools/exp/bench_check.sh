#!/bin/bash
# the default bench line and the --mode ref line of a small cube: are the JSON lines well formed, is parity green
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2last_bench.json 2> gpurun_out/r2last_bench.err; tail -2 gpurun_out/r2last_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2last_bench.json"))
print("headline", round(d["ms_per_step"], 2), "ms e2e", round(d["e2e"]["ms_per_step"], 2), "parity", d["parity"]["mismatches"], "/", d["parity"]["checked_entries"], "launches", d["gpu_launches"])
print(json.dumps(d["extra"]["qgemm_reference_order"])[:1500])
print("errors:", [k for k in d["extra"] if "error" in k.lower()])
PY
timeout 300 python bench.py --mode ref --shape 2048,2048,2048 --steps 3 --warmup 3 > gpurun_out/r2last_bench_ref.json 2> gpurun_out/r2last_bench_ref.err; tail -2 gpurun_out/r2last_bench_ref.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2last_bench_ref.json"))
print("ref-mode headline", d["value"], d["ms_per_step"], json.dumps(d["roofline"])[:900], "parity", d.get("parity"))
PY
