#!/bin/bash
# ncu --set full of the BVH kernels of the third wave iteration of a bistro render; per-source-line summaries only
mkdir -p gpurun_out
export PROF_RR_DELTA=0.05
timeout 300 python tools/prof_run.py bistro 2 > gpurun_out/r2_bvh_plain.log 2>&1 || { echo plain run failed; tail -5 gpurun_out/r2_bvh_plain.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name 'regex:k_occl_bvh|k_closest_bvh|k_closest_finish' --launch-skip 6 --launch-count 3 -f -o /tmp/bvh python tools/prof_run.py bistro 2 > gpurun_out/r2_bvh_ncu.log 2>&1
for i in 0 1 2; do
  ncu -i /tmp/bvh.ncu-rep --page source --csv --print-source sass,cuda --launch-skip $i --launch-count 1 > /tmp/bvh_$i.csv 2>/dev/null
  python tools/ncu_lines.py /tmp/bvh_$i.csv 60 > gpurun_out/r2_lines_bvh_$i.txt 2>&1
done
tools/ncu_summary.sh /tmp/bvh.ncu-rep gpurun_out/r2_ncu_bvh.txt k_occl_bvh k_closest_bvh k_closest_finish
