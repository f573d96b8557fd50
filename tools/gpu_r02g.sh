#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02g}
timeout 1500 python -m pytest tests -m gpu -q -s > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log
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
    env $HOPENV HOP_FUSED_PROFILE=1 timeout 600 python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline "$@" 2>&1 | grep "fused profile" | tail -1
  done
}
HOPENV="HOP_X=0" run default
HOPENV="HOP_FUSED_VARIANT=2" run v2
HOPENV="HOP_FUSED_VARIANT=1" run v1
HOPENV="HOP_FUSED_SLOTS=1" run slots1
HOPENV="HOP_FUSED_SLOTS=2" run slots2
HOPENV="HOP_FUSED_SLOTS=8" run slots8
