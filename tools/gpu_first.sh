#!/bin/bash
# first GPU contact: parity tests, then a tiny render timing
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
make -s -C oracle liblumo_oracle.so
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.txt
cat gpurun_out/pytest_gpu.txt
