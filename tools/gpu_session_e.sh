#!/bin/bash
# ncu --set full captures of the kernels of the fast-mode qgemm (one unit launch each) and of k_gemm
set -u
TAG=${1:-r2e}
mkdir -p gpurun_out
cap() { # name regex skip driver...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c 1 -f -o gpurun_out/${TAG}_${name} "$@" > gpurun_out/${TAG}_${name}.log 2>&1
  python profiles/summarize.py rep gpurun_out/${TAG}_${name}.ncu-rep gpurun_out/${TAG}_${name}_ncu_full.txt && grep -E "kernel:|gpu__time_duration.sum|pipe_tensor_cycles_active.avg.pct|dram__bytes_(read|write).sum \[|sm__issue_active.avg.pct|registers_per_thread|warps_active.avg.pct" gpurun_out/${TAG}_${name}_ncu_full.txt | head -12
}
cap mma "k_oz_mma" 20 python tools/ncu_qgemm.py 8192 D113 2
cap fold "k_crt_fold" 20 python tools/ncu_qgemm.py 8192 D113 2
cap resA "k_crt_residues<" 5 python tools/ncu_qgemm.py 8192 D113 2
cap resB "k_crt_residues_t" 5 python tools/ncu_qgemm.py 8192 D113 2
cap fold_dexp "k_crt_fold" 20 python tools/ncu_qgemm.py 8192 Dexp 2
cap kgemm "k_gemm" 1 python tools/ncu_kgemm.py 2048
ls -la gpurun_out/${TAG}_*.ncu-rep
