#!/bin/bash
mkdir -p gpurun_out
PROF_WARM=0 python tools/prof_run.py bunny 2 > gpurun_out/prof6_plain.log 2>&1 &&
PROF_WARM=0 ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum --clock-control none -s 20 -c 120 --csv --log-file gpurun_out/launches_nm.csv python tools/prof_run.py bunny 2 > gpurun_out/ncu_launch6.log 2>&1
tail -2 gpurun_out/ncu_launch6.log
