#!/bin/bash
# ncu captures of the final build: k_gemm_nb (full set) and the fast qdot kernel durations against n (launch list)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_gemm_nb" -s 1 -c 1 -f -o gpurun_out/r2last_kgemm_nb python tools/ncu_kgemm.py 2048 > gpurun_out/r2last_kgemm_nb.log 2>&1
python profiles/summarize.py rep gpurun_out/r2last_kgemm_nb.ncu-rep gpurun_out/r2last_kgemm_nb_ncu_full.txt && grep -E "gpu__time_duration.sum|issue_active.avg.pct_of_peak_sustained_active|pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed|fmaheavy|smsp__inst_executed.sum" gpurun_out/r2last_kgemm_nb_ncu_full.txt
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_dot" --csv --log-file gpurun_out/r2last_dot_sizes.csv python tools/exp/dot_sizes.py ncu > /dev/null 2>&1
grep -v "^==" gpurun_out/r2last_dot_sizes.csv | awk -F'","' '{print $5, $NF}' | tail -7
