#!/bin/bash
# Session for the two kernels added late in round 2: k_gemm_nb (reference-order qgemm) and k_dot_wide_tma (fast qdot): side-by-side
# timing, every GPU test, ncu --set full of one launch of each.  usage: gpurun --timeout 1500 -- 'bash tools/gpu_session_r2z.sh TAG'
set -u
TAG=${1:-r2z}
mkdir -p gpurun_out
echo "== A/B"; python tools/exp/kgemm_ab.py 2048 D113 2>&1 | tail -2; python tools/exp/kgemm_ab.py 4096 D113 2>&1 | tail -2
echo "== all gpu tests"; timeout 1200 python -m pytest tests -m gpu -q --timeout 600 2>&1 | grep -v "^    \|^$" | tail -12 | tee gpurun_out/${TAG}_pytest_all.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_gemm_nb" -s 1 -c 1 -f -o gpurun_out/${TAG}_kgemm_nb python tools/ncu_kgemm.py 2048 > gpurun_out/${TAG}_kgemm_nb.log 2>&1
python profiles/summarize.py rep gpurun_out/${TAG}_kgemm_nb.ncu-rep gpurun_out/${TAG}_kgemm_nb_ncu_full.txt && grep -E "gpu__time_duration.sum|issue_active.avg.pct_of_peak_sustained_active|pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed|fmaheavy|smsp__inst_executed.sum" gpurun_out/${TAG}_kgemm_nb_ncu_full.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_dot_wide_tma" -s 1 -c 1 -f -o gpurun_out/${TAG}_dot_tma python tools/exp/dot_ncu.py > gpurun_out/${TAG}_dot_tma.log 2>&1
python profiles/summarize.py rep gpurun_out/${TAG}_dot_tma.ncu-rep gpurun_out/${TAG}_dot_tma_ncu_full.txt && grep -E "gpu__time_duration.sum|issue_active.avg.pct_of_peak_sustained_active|dram__bytes_read.sum \[|gpu__dram_throughput" gpurun_out/${TAG}_dot_tma_ncu_full.txt
timeout 200 python tools/exp/dot_bench.py 2>&1 | grep "variant 1" | tail -4
