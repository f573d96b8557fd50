#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r03i}
run() { local name=$1; shift; local wl=$1; shift
  env "$@" timeout 600 python bench.py --workload $wl --steps 8 --warmup 3 --no-also --no-stages --no-cpu-baseline > $OUT/${TAG}_${name}.json 2> $OUT/${TAG}_${name}.err
  python - <<PY
import json
try:
    d=json.loads(open('$OUT/${TAG}_${name}.json').read().strip().splitlines()[-1])
    k=d['config']['kernel_ms']
    print('$name', 'value %.4g ms/step %.4f icp %.4f lcp %.4f nn_build %.4f'%(d['value'],d['ms_per_step'],d['roofline']['kernel_ms_per_launch'],k['lcp_score']['ms_per_step'],k['nn_build']['ms_per_step']), d['config']['nn_grid_icp'], d['config']['nn_grid_lcp'])
except Exception as e:
    print('$name', 'no line', e); print(open('$OUT/${TAG}_${name}.err').read()[-400:])
PY
}
for sc in 0.5 0.7 0.85 1.0 1.25 1.6 2.0; do
  run headline_vs$sc headline HOP_VOXEL_SCALE=$sc
done
for sc in 0.7 1.0 1.4; do
  run C2_vs$sc C2 HOP_VOXEL_SCALE=$sc
done
