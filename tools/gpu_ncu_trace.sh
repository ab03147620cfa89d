#!/bin/bash
# ncu --set full of the wave traversal kernels, iterations 1-3 (run only after the same command exited 0 without ncu)
mkdir -p gpurun_out
PROF_WARM=0 python tools/prof_run.py bunny 4 > gpurun_out/prof_trace_plain.log 2>&1 &&
PROF_WARM=0 ncu --set full --clock-control none --import-source on -k regex:"k_wave_trace|k_wave_occlude" -s 2 -c 6 -f -o gpurun_out/prof_trace python tools/prof_run.py bunny 4 > gpurun_out/ncu_trace.log 2>&1
tail -1 gpurun_out/ncu_trace.log | cut -c1-200
