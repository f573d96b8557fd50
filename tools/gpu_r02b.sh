#!/bin/bash
# round 2, visit b: slots + cold-path split; variants
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02b}
timeout 1500 python -m pytest tests -m gpu -q -s > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log
grep -E "passed|failed|FAILED" $OUT/${TAG}_pytest_gpu.log | tail -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?" >> $OUT/${TAG}_smoke.log; tail -2 $OUT/${TAG}_smoke.log
run() { # name, env...
  local NAME=$1; shift
  for WL in C2 headline; do
    env "$@" timeout 900 python bench.py --workload $WL --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_${WL}_${NAME}.json 2> $OUT/${TAG}_bench_${WL}_${NAME}.err
    python - <<PY
import json
d=json.loads(open('$OUT/${TAG}_bench_${WL}_${NAME}.json').read().strip().splitlines()[-1])
print('$NAME $WL', 'value %.4g ms/step %.4f e2e %.4g kernel_ms %.4f frac %.3f'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['kernel_ms_per_launch'],d['roofline']['frac']))
PY
  done
}
run default HOP_X=0
run v1 HOP_FUSED_VARIANT=1
run v2 HOP_FUSED_VARIANT=2
run slots1 HOP_FUSED_SLOTS=1
run slots2 HOP_FUSED_SLOTS=2
run slots4 HOP_FUSED_SLOTS=4
for WL in C2 headline; do HOP_FUSED_PROFILE=1 timeout 600 python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | grep "fused profile" | tail -1; done
