#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02e}
timeout 1500 python -m pytest tests -m gpu -q -s -x > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log
grep -E "passed|failed|FAILED" $OUT/${TAG}_pytest_gpu.log | tail -20
run() {
  local NAME=$1; shift
  for WL in C2 headline; do
    env $HOPENV timeout 900 python bench.py --workload $WL --steps 10 --warmup 3 --no-cpu-baseline "$@" > $OUT/${TAG}_bench_${WL}_${NAME}.json 2> $OUT/${TAG}_bench_${WL}_${NAME}.err
    python - <<PY
import json
d=json.loads(open('$OUT/${TAG}_bench_${WL}_${NAME}.json').read().strip().splitlines()[-1])
k=d['config']['kernel_ms']
print('$NAME $WL', 'value %.4g ms/step %.4f e2e %.4g frac %.3f'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac']), {a:round(b['ms_per_step'],4) for a,b in k.items()})
PY
  done
}
HOPENV="HOP_X=0" run default
HOPENV="HOP_X=0" run sync --pipeline 1
HOPENV="HOP_FUSED_VARIANT=2" run v2
HOPENV="HOP_FUSED_VARIANT=1" run v1
HOPENV="HOP_FUSED_SLOTS=1" run slots1
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio --clock-control none -k regex:"icp_solve" -c 4 --csv --log-file $OUT/${TAG}_solve_launches.csv python bench.py --workload C2 --steps 1 --warmup 3 --no-cpu-baseline --pipeline 1 > $OUT/${TAG}_ncu.log 2>&1
grep -v "^==" $OUT/${TAG}_solve_launches.csv | awk -F'","' '{print $1, $5, $(NF-2), $NF}' | head -20
