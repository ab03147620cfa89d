#!/bin/bash
# weak-scaling line at N GPUs: tools/gpu_scale.sh N
N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 --other-scenes '' --no-cpu > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -2 gpurun_out/bench_n$N.err | cut -c1-300
cut -c1-400 gpurun_out/bench_n$N.json
