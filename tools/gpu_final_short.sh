#!/bin/bash
# final check of a round on one GPU: -m gpu suite, bistro launch list (time + DRAM bytes) -> profiles/traffic.json, full bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_final_gpu_tests.log 2>&1; echo "gpu suite rc=$?"; tail -3 gpurun_out/r2_final_gpu_tests.log
export PROF_RR_DELTA=0.05
for WS in "bistro 2" "bunny 8"; do
  set -- $WS; W=$1; SPP=$2
  PROF_ITERLOG=gpurun_out/r2_iterlog_$W.json timeout 300 python tools/prof_run.py $W $SPP > gpurun_out/r2_prof_$W.log 2>&1 || { echo "plain run of $W failed"; continue; }
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches_${W}_final.csv python tools/prof_run.py $W $SPP > /dev/null 2>&1
  python tools/traffic_from_ncu.py gpurun_out/r2_launches_${W}_final.csv gpurun_out/r2_iterlog_$W.json $W profiles/traffic.json > /dev/null 2> gpurun_out/r2_traffic_$W.err || tail -3 gpurun_out/r2_traffic_$W.err
done
cp profiles/traffic.json gpurun_out/traffic.json
timeout 900 python bench.py > gpurun_out/r2_bench_last.json 2> gpurun_out/r2_bench_last.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_bench_last.err | cut -c1-300; cut -c1-400 gpurun_out/r2_bench_last.json
