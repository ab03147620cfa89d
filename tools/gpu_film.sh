#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_film_encode.py -m gpu -x -q 2>&1 | tail -6
timeout 100 python tools/film_run.py 5 2>&1 | tail -4
LUMO_FILM_F64=1 timeout 100 python tools/film_run.py 4 2>&1 | tail -2
timeout 200 ncu --set full --clock-control none -k regex:k_film_encode -c 2 -o gpurun_out/film_encode python tools/film_run.py 2 > gpurun_out/film_ncu.log 2>&1; tail -2 gpurun_out/film_ncu.log
