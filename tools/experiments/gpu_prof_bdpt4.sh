#!/bin/bash
mkdir -p gpurun_out
for c in 131072 1048576; do
LUMO_BDPT_BATCH=$c PROF_WARM=0 ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_bdpt_connect --csv --log-file gpurun_out/launches_bdptc_$c.csv python tools/prof_run.py caustics_bdpt 1 > gpurun_out/ncu_bdpt4.log 2>&1
done
tail -1 gpurun_out/ncu_bdpt4.log | cut -c1-200
