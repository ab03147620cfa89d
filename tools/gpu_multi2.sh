#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | cut -c1-20
timeout 200 python -m pytest tests/test_render_multi.py -m gpu -x -q 2>&1 | tail -6
timeout 200 python tools/multi_run.py bunny 8 2> gpurun_out/multi_run.err | tee gpurun_out/multi_run.jsonl; tail -3 gpurun_out/multi_run.err
