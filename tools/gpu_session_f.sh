#!/bin/bash
# Final single-GPU session of a round: every GPU test, the bench line under the driver's protocol, ncu --set full of one tensor-kernel
# launch of the default unit shape (-> profiles/ncu_traffic.json), launch list.  usage: gpurun --timeout 3000 -- 'bash tools/gpu_session_f.sh TAG'
set -u
TAG=${1:-r2f}
mkdir -p gpurun_out
echo "== all gpu tests"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | grep -v "^    \|^$" | tail -30 | tee gpurun_out/${TAG}_pytest_all.log
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench.json"))
    r = d["roofline"]
    print("headline", round(d["ms_per_step"], 3), "ms", round(d["value"] / 1e3, 2), "TFLOP/s; mma", round(r["kernel_ms"], 2), "ms frac", round(r["frac"], 3), "whole", round(r["whole_step_frac"], 3), "e2e", round(d["e2e"]["ms_per_step"], 2), "ms; parity", d["parity"]["mismatches"], d["parity"]["checked_entries"], "clocks", d["clocks"], "traffic", r["traffic"])
    for k, v in d["extra"].items():
        print(k, json.dumps(v)[:330])
except Exception as e:
    print("bench parse failed", e)
PY
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-400
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_oz_mma" -s 12 -c 1 -f -o gpurun_out/${TAG}_mma python tools/ncu_qgemm.py 8192 D113 2 > gpurun_out/${TAG}_mma.log 2>&1
python profiles/summarize.py rep gpurun_out/${TAG}_mma.ncu-rep gpurun_out/${TAG}_mma_ncu_full.txt && grep -E "gpu__time_duration.sum|pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed|dram__bytes_(read|write).sum \[|launch__grid_size" gpurun_out/${TAG}_mma_ncu_full.txt; tail -2 gpurun_out/${TAG}_mma.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_|qb" -c 120 --csv --log-file gpurun_out/${TAG}_launches.csv python tools/ncu_qgemm.py 8192 D113 2 > /dev/null 2>&1
python profiles/summarize.py launches gpurun_out/${TAG}_launches.csv gpurun_out/${TAG}_launches.txt && tail -8 gpurun_out/${TAG}_launches.txt
