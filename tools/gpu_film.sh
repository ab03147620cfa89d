#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_film_encode.py -m gpu -x -q 2>&1 | tail -3
timeout 100 python tools/film_run.py 5 2>&1 | tail -5
timeout 200 ncu --set full --clock-control none -k regex:k_film_encode -c 2 -o gpurun_out/film_encode python tools/film_run.py 2 > gpurun_out/film_ncu.log 2>&1; tail -2 gpurun_out/film_ncu.log
timeout 200 python bench.py --other-scenes "" --no-cpu > gpurun_out/bench_film2.json 2> gpurun_out/bench_film2.err; echo "bench rc=$?"
python -c "import json; d=json.load(open('gpurun_out/bench_film2.json')); print(d['value'], d['film_encode'])"
