#!/bin/bash
for cfg in "0 bunny 4" "1 bunny 2" "1 bunny 4" "0 cornell 8" "1 cornell 8"; do
  set -- $cfg
  echo "WARM=$1 $2 $3"; LUMO_TRACE_NM=1 PROF_WARM=$1 timeout 300 python tools/prof_run.py $2 $3 2>&1 | tail -1 | cut -c1-420
done
