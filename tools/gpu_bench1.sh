#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err
tail -5 gpurun_out/bench1.err
cat gpurun_out/bench1.json
