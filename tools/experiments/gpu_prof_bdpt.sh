#!/bin/bash
mkdir -p gpurun_out
PROF_WARM=0 python tools/prof_run.py caustics_bdpt 1 > gpurun_out/prof_bdpt_plain.log 2>&1 &&
PROF_WARM=0 ncu --set full --clock-control none --import-source on -k regex:"k_bdpt_connect|k_bdpt_walk" -s 4 -c 2 -f -o gpurun_out/prof_bdpt python tools/prof_run.py caustics_bdpt 1 > gpurun_out/ncu_bdpt.log 2>&1
tail -1 gpurun_out/ncu_bdpt.log | cut -c1-300
