#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 300 python -m pytest tests/test_render_multi.py tests/test_film_encode.py -m gpu -x -q 2>&1 | tail -12
timeout 100 python tools/film_run.py 5 2>&1 | tail -4
