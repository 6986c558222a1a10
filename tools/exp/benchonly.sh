TAG=r2final
SECONDS=0
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -3 gpurun_out/${TAG}_bench.err; echo "bench wall ${SECONDS}s"
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench.json"))
r = d["roofline"]
print("headline", round(d["ms_per_step"], 3), "ms", round(d["value"] / 1e3, 2), "TFLOP/s; mma", round(r["kernel_ms"], 2), "ms frac", round(r["frac"], 3), "whole", round(r["whole_step_frac"], 3), "e2e", round(d["e2e"]["ms_per_step"], 2), "ms; parity", d["parity"]["mismatches"], d["parity"]["checked_entries"], "clocks", d["clocks"])
for k, v in d["extra"].items():
    print(k, json.dumps(v)[:260])
PY
