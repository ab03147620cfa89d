#!/bin/bash
# quick perf check: bunny + bistro short renders, then the GPU test-suite
mkdir -p gpurun_out
python tools/prof_run.py bunny 4 2>&1 | tail -1
python tools/prof_run.py bistro 1 2>&1 | tail -1
python tools/prof_run.py conference 4 2>&1 | tail -1
make -s -C oracle liblumo_oracle.so 2>/dev/null
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
