#!/bin/bash
# Level-1/2 session: ncu --set full of the sliced FP64 kernels (k_gemv_f64 both layouts, k_sumsq_f64), their launch list, and the
# variant comparison.  usage: gpurun --timeout 1500 -- 'bash tools/gpu_session_q.sh TAG'
set -u
TAG=${1:-r2q}
mkdir -p gpurun_out
for lay in R C; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"^k_gemv_f64$" -s 1 -c 1 -f -o gpurun_out/${TAG}_gemv_${lay} python tools/exp/gemv_ncu.py 32768 2 $lay > gpurun_out/${TAG}_gemv_${lay}.log 2>&1
  python profiles/summarize.py rep gpurun_out/${TAG}_gemv_${lay}.ncu-rep gpurun_out/${TAG}_gemv_${lay}_ncu_full.txt && grep -E "gpu__time_duration.sum|issue_active.avg.pct_of_peak_sustained_active|dram__bytes_read.sum \[" gpurun_out/${TAG}_gemv_${lay}_ncu_full.txt
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_sumsq_f64" -s 1 -c 1 -f -o gpurun_out/${TAG}_sumsq python tools/exp/nrm2_ncu.py 100000000 > gpurun_out/${TAG}_sumsq.log 2>&1
python profiles/summarize.py rep gpurun_out/${TAG}_sumsq.ncu-rep gpurun_out/${TAG}_sumsq_ncu_full.txt && grep -E "gpu__time_duration.sum|issue_active.avg.pct_of_peak_sustained_active|dram__bytes_read.sum \[" gpurun_out/${TAG}_sumsq_ncu_full.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_|qb" -c 40 --csv --log-file gpurun_out/${TAG}_gemv_launches.csv python tools/exp/gemv_ncu.py 32768 2 R > /dev/null 2>&1
python profiles/summarize.py launches gpurun_out/${TAG}_gemv_launches.csv gpurun_out/${TAG}_gemv_launches.txt && tail -12 gpurun_out/${TAG}_gemv_launches.txt
timeout 200 python tools/exp/gemv_bench.py 2>&1 | tail -8
timeout 200 python tools/exp/nrm2_bench.py 2>&1 | tail -8
