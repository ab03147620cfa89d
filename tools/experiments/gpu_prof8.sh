#!/bin/bash
mkdir -p gpurun_out
LUMO_TRACE_NM=1 PROF_WARM=1 ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum --clock-control none -s 330 -c 70 --csv --log-file gpurun_out/launches_nm3.csv python tools/prof_run.py bunny 4 > gpurun_out/ncu_launch8.log 2>&1
tail -2 gpurun_out/ncu_launch8.log | cut -c1-600
