#!/bin/bash
mkdir -p gpurun_out
PROF_WARM=0 ncu --section SpeedOfLight --section WarpStateStats --section SourceCounters --section Occupancy --section MemoryWorkloadAnalysis --clock-control none --import-source on -k regex:"k_bdpt_connect" -s 2 -c 1 -f -o gpurun_out/prof_bdpt2 python tools/prof_run.py caustics_bdpt 1 > gpurun_out/ncu_bdpt2.log 2>&1
tail -1 gpurun_out/ncu_bdpt2.log | cut -c1-300
