#!/bin/bash
mkdir -p gpurun_out
PROF_WARM=0 ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum --clock-control none -s 60 -c 320 --csv --log-file gpurun_out/launches_bdpt.csv python tools/prof_run.py caustics_bdpt 1 > gpurun_out/ncu_bdpt3.log 2>&1
tail -1 gpurun_out/ncu_bdpt3.log | cut -c1-200
