#!/bin/bash
mkdir -p gpurun_out
W=${1:-cornell}
LUMO_TRACE_NM=1 PROF_WARM=0 python tools/prof_run.py $W 2 > gpurun_out/prof7_plain.log 2>&1 &&
LUMO_TRACE_NM=1 PROF_WARM=0 ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/launches_nm2.csv python tools/prof_run.py $W 2 > gpurun_out/ncu_launch7.log 2>&1
tail -2 gpurun_out/ncu_launch7.log
