#!/bin/bash
# round-end check: film-encode tests first (new code), then the whole -m gpu suite, then one bench line
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_film_encode.py -m gpu -x -q > gpurun_out/final_film_tests.log 2>&1; echo "film tests rc=$?"; tail -4 gpurun_out/final_film_tests.log
timeout 430 python -m pytest tests -m gpu -x -q --durations=12 --deselect tests/test_film_encode.py > gpurun_out/final_gpu_tests.log 2>&1; echo "gpu suite rc=$?"; tail -18 gpurun_out/final_gpu_tests.log
timeout 170 python bench.py --other-scenes "" > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/final_bench.err | cut -c1-300; cut -c1-600 gpurun_out/final_bench.json
