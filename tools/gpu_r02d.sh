#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02d}
timeout 600 ncu --metrics gpu__time_duration.sum,launch__registers_per_thread,launch__grid_size,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,smsp__average_warp_latency_issue_stalled_no_instruction.pct,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio --clock-control none -k regex:"icp_solve|icp_moments" -c 24 --csv --log-file $OUT/${TAG}_solve_launches.csv python bench.py --workload C2 --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02d_solve_launches.csv')) if len(r)>10]
hdr=rows[0]; 
ik=hdr.index('Kernel Name'); im=hdr.index('Metric Name'); iv=hdr.index('Metric Value'); iid=hdr.index('ID')
cur={}
for r in rows[1:]:
    cur.setdefault(r[iid],{'k':r[ik][:40]})[r[im]]=r[iv]
for i,d in cur.items(): print(i,d)
PY
