#!/bin/bash
mkdir -p gpurun_out
make -s -C oracle liblumo_oracle.so
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -150 > gpurun_out/pytest_gpu.txt
tail -15 gpurun_out/pytest_gpu.txt
