#!/bin/bash
# full measurement pass: bench (ours + reference arm), then the ncu launch list of the same profiling command
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
tail -3 gpurun_out/bench_full.err
python bench.py --impl reference > gpurun_out/bench_ref.json 2>> gpurun_out/bench_full.err
python tools/prof_run.py bunny 4 > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_bunny.csv python tools/prof_run.py bunny 4 > gpurun_out/ncu_launch.log 2>&1
tail -1 gpurun_out/ncu_launch.log | cut -c1-200
