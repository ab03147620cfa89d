#!/bin/bash
# ncu --set full of the NEE kernels of the third wave iteration of a bistro render; per-source-line summaries only (the report stays on the box)
mkdir -p gpurun_out
export PROF_RR_DELTA=0.05
timeout 300 python tools/prof_run.py bistro 2 > gpurun_out/r2_nee_plain.log 2>&1 || { echo plain run failed; tail -5 gpurun_out/r2_nee_plain.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name 'regex:k_nee_(a|b|eval)' --launch-skip 18 --launch-count 3 -f -o /tmp/nee python tools/prof_run.py bistro 2 > gpurun_out/r2_nee_ncu.log 2>&1
for i in 0 1 2; do
  ncu -i /tmp/nee.ncu-rep --page source --csv --print-source sass,cuda --launch-skip $i --launch-count 1 > /tmp/nee_$i.csv 2>/dev/null
  python tools/ncu_lines.py /tmp/nee_$i.csv 70 > gpurun_out/r2_lines_nee_$i.txt 2>&1
done
tools/ncu_summary.sh /tmp/nee.ncu-rep gpurun_out/r2_ncu_nee.txt k_nee_a k_nee_b k_nee_eval
