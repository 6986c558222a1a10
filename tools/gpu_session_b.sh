#!/bin/bash
set -u
TAG=${1:-r2b}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_hostpath.py -q --timeout 300 -k "cancelling or inf_nan or streamed_host" 2>&1 | grep -v "^    \|^$" | tail -120 > gpurun_out/${TAG}_pytest.log; tail -60 gpurun_out/${TAG}_pytest.log
timeout 600 python tools/tune_units.py 8192 D113 2>&1 | tee gpurun_out/${TAG}_tune_D113.log
timeout 300 python tools/tune_units.py 8192 Dexp 2048x2048,2048x8192 2>&1 | tee gpurun_out/${TAG}_tune_Dexp.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_|qb" -c 160 --csv --log-file gpurun_out/${TAG}_launches_dexp.csv python tools/ncu_qgemm.py 8192 Dexp 2 > /dev/null 2>&1
python profiles/summarize.py launches gpurun_out/${TAG}_launches_dexp.csv gpurun_out/${TAG}_launches_dexp.txt && tail -12 gpurun_out/${TAG}_launches_dexp.txt
