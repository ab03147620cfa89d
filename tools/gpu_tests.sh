#!/bin/bash
mkdir -p gpurun_out
make -s -C oracle liblumo_oracle.so 2>/dev/null
timeout 1700 python -m pytest tests -m gpu -q -x "$@" 2>&1 | tail -15
