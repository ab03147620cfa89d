#!/bin/bash
# node-major v2 (inline light objects + lane-refilled heavy kd kernel) vs the nested default; then the GPU test-suite
mkdir -p gpurun_out
for s in "bunny 4" "cornell 8" "dragon 2" "conference 4"; do
  echo "NM=1 $s"; LUMO_TRACE_NM=1 timeout 300 python tools/prof_run.py $s 2>&1 | tail -1 | cut -c1-400
  echo "NM=0 $s"; LUMO_TRACE_NM=0 timeout 300 python tools/prof_run.py $s 2>&1 | tail -1 | cut -c1-400
done
make -s -C oracle liblumo_oracle.so 2>/dev/null
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
