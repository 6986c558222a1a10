#!/bin/bash
# First GPU call of the next session: validates what was added after the last GPU run of round 1 and measures the opt-in
# row-pass shape.  Run from the repo root under gpurun, e.g.
#   gpurun --timeout 600 -- 'bash tools/next_gpu_session.sh'
# Outputs land in gpurun_out/ (copy what should be judged into profiles/).
set -u
mkdir -p gpurun_out
# 1. the GPU tests added without a GPU run (pre-checked with the CPU build of the same arithmetic)
timeout 300 python -m pytest tests/test_gpu_ozaki.py -x -q -k "other_word_and_group_counts or pass_shapes_agree or peer_outputs" 2>&1 | tail -5
# 2. equal row passes (default) vs short first / last pass, same box, back to back
for shape in 0 1; do
  timeout 200 python bench.py --steps 5 --warmup 3 --no-extra --pass-shape $shape > gpurun_out/bench_pass_shape_$shape.json 2> gpurun_out/bench_pass_shape_$shape.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_pass_shape_$shape.json"))
print("pass shape $shape:", round(d["ms_per_step"], 3), "ms,", round(d["value"] / 1e3, 2), "TFLOP/s, passes", d["roofline"]["plan"]["row_passes"], "mma ms", round(d["roofline"]["kernel_ms"], 2))
PY
done
# 2b. e2e with more C slabs in the pipelined host path (shorter tail after the last upload)
for slabs in 4 8; do
  timeout 200 python bench.py --steps 3 --warmup 3 --no-extra --host-slabs $slabs > gpurun_out/bench_host_slabs_$slabs.json 2> /dev/null
  python -c "import json; d = json.load(open('gpurun_out/bench_host_slabs_$slabs.json')); print('host slabs $slabs: e2e', round(d['e2e']['ms_per_step'], 2), 'ms')"
done
# 3. launch list of one call per shape (shares of the summed kernel time)
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_|qb" -c 80 --csv --log-file gpurun_out/launches_residues.csv python tools/ncu_qgemm.py 8192 residues 2 > /dev/null 2>&1
python profiles/summarize.py launches gpurun_out/launches_residues.csv gpurun_out/launches_residues.txt && cat gpurun_out/launches_residues.txt
