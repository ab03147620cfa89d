#!/bin/bash
mkdir -p gpurun_out
export LUMO_GPU_SO=$PWD/lumo_b200/liblumo_gpu_k4.so
LUMO_TRACE_NM=1 PROF_WARM=0 ncu --set full --clock-control none --import-source on -k regex:"k_nm_heavy" -s 4 -c 2 -f -o gpurun_out/prof9 python tools/prof_run.py bunny 2 > gpurun_out/ncu_full9.log 2>&1
tail -1 gpurun_out/ncu_full9.log | cut -c1-300
