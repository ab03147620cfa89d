#!/bin/bash
mkdir -p gpurun_out
python tools/prof_run.py bunny 2 > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bunny.csv python tools/prof_run.py bunny 2 > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_wave_ -s 4 -c 6 -o gpurun_out/prof_bunny python tools/prof_run.py bunny 2 > gpurun_out/ncu_full.log 2>&1
cat gpurun_out/prof_plain.log; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out
