#!/bin/bash
set -u
TAG=${1:-r2d}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_hostpath.py -q --timeout 300 2>&1 | grep -v "^    \|^$" | tail -40 > gpurun_out/${TAG}_pytest.log; tail -6 gpurun_out/${TAG}_pytest.log
for prio in 0121 0111 0211 0122; do
  echo "== QBLAS_STREAM_PRIO=$prio"
  QBLAS_STREAM_PRIO=$prio timeout 600 python tools/tune_units.py 8192 D113 2048x2048,2048x4096 0x0,512x1024,1024x1024,512x2048 2>&1 | grep "^unit\|gaps" | tee -a gpurun_out/${TAG}_tune.log
done
