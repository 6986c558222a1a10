#!/bin/bash
set -u
TAG=${1:-r2c}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_hostpath.py -q --timeout 300 2>&1 | grep -v "^    \|^$" | tail -40 > gpurun_out/${TAG}_pytest.log; tail -15 gpurun_out/${TAG}_pytest.log
timeout 600 python tools/tune_units.py 8192 D113 2048x2048,2048x4096,2048x8192,2048x2304,4096x2048 2>&1 | tee gpurun_out/${TAG}_tune_D113.log
timeout 900 python tools/tune_rect.py 2>&1 | tee gpurun_out/${TAG}_tune_rect.log
