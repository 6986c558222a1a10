set -x
python tools/exp/kgemm_ab.py 2048 D113 2>&1 | tail -2
python tools/exp/kgemm_ab.py 1024 Dexp 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_fast_level12.py -x -q -m gpu 2>&1 | tail -5
