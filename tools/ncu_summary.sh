#!/bin/bash
# Runs on the GPU box after an `ncu --set full -o <rep>` capture: keeps text summaries, drops the (large) report.
# usage: tools/ncu_summary.sh <report.ncu-rep> <out.txt> <kernel regex> [<kernel regex> ...]
REP=$1; OUT=$2; shift 2
{
  echo "# ncu --set full summary of $REP"
  ncu -i "$REP" --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr=rows[0]; ix={h:i for i,h in enumerate(hdr)}
keys=['gpu__time_duration.sum','launch__registers_per_thread','launch__grid_size','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__t_bytes.sum']
for r in rows[2:]:
    print(r[ix['Kernel Name']][:90])
    for k in keys:
        if k in ix: print('    %-72s %s %s'%(k, r[ix[k]], rows[1][ix[k]]))
"
  for K in "$@"; do echo; python tools/ncu_stalls.py "$REP" "$K" 0; done
} > "$OUT" 2>&1
rm -f "$REP"
