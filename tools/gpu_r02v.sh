#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02v}
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/${TAG}_pytest.log
run() { local name=$1; shift; local wl=$1; shift
  env "$@" timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-also --no-stages --no-cpu-baseline > $OUT/${TAG}_${name}.json 2> $OUT/${TAG}_${name}.err
  python - <<PY
import json
try:
    d=json.loads(open('$OUT/${TAG}_${name}.json').read().strip().splitlines()[-1])
    k=d['config']['kernel_ms']
    print('$name', 'value %.4g ms/step %.4f icp %.4f lcp %.4f nn_build %.4f'%(d['value'],d['ms_per_step'],d['roofline']['kernel_ms_per_launch'],k['lcp_score']['ms_per_step'],k['nn_build']['ms_per_step']), d['config']['nn_grid_icp'])
except Exception as e:
    print('$name', 'no line', e); print(open('$OUT/${TAG}_${name}.err').read()[-600:])
PY
  grep "hop fused profile" $OUT/${TAG}_${name}.err | tail -1
}
run headline headline HOP_X=0
run headline_prof headline HOP_FUSED_PROFILE=1
run C2 C2 HOP_X=0
run C3 C3 HOP_X=0
run C4 C4 HOP_X=0
run C5 C5 HOP_X=0
