#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02s}
run() { local name=$1; shift; local wl=$1; shift
  env "$@" timeout 600 python bench.py --workload $wl --steps 8 --warmup 3 --no-also --no-stages --no-cpu-baseline > $OUT/${TAG}_${name}.json 2> $OUT/${TAG}_${name}.err
  python - <<PY
import json
try:
    d=json.loads(open('$OUT/${TAG}_${name}.json').read().strip().splitlines()[-1])
    print('$name', 'value %.4g ms/step %.4f icp %.4f'%(d['value'],d['ms_per_step'],d['roofline']['kernel_ms_per_launch']))
except Exception as e:
    print('$name', 'no line', e); print(open('$OUT/${TAG}_${name}.err').read()[-300:])
PY
}
for wl in C2 C4 C3 headline; do
  for v in 1 2 3; do
    for s in 0 1 2; do
      run ${wl}_v${v}_s${s} $wl HOP_FUSED_VARIANT=$v HOP_FUSED_SLOTS=$s
    done
  done
done
