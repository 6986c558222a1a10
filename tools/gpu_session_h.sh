#!/bin/bash
set -u
TAG=${1:-r2j}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_level12.py tests/test_gpu_fast_level12.py tests/test_gpu_hostpath.py tests/test_gpu_gemm.py -q --timeout 600 2>&1 | grep -v "^    \|^$" | tail -30 | tee gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
