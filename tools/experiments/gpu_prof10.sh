#!/bin/bash
mkdir -p gpurun_out
PROF_WARM=0 ncu --set full --clock-control none --import-source on -k regex:"k_nee|k_scatter" -s 4 -c 4 -f -o gpurun_out/prof10 python tools/prof_run.py ${1:-bunny} 2 > gpurun_out/ncu_full10.log 2>&1
tail -1 gpurun_out/ncu_full10.log | cut -c1-300
