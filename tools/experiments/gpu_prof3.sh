#!/bin/bash
mkdir -p gpurun_out
python tools/prof_run.py bunny 2 > gpurun_out/prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_wave_trace|k_wave_occlude" -s 6 -c 4 -o gpurun_out/prof_bunny3 python tools/prof_run.py bunny 2 > gpurun_out/ncu_full.log 2>&1
cat gpurun_out/prof_plain.log; tail -2 gpurun_out/ncu_full.log
