#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r03s}
run() { local name=$1; shift; local wl=$1; shift
  env "$@" timeout 600 python bench.py --workload $wl --steps 6 --warmup 3 --no-also --no-stages --no-cpu-baseline > $OUT/${TAG}_${name}.json 2> $OUT/${TAG}_${name}.err
  python - <<PY
import json
try:
    d=json.loads(open('$OUT/${TAG}_${name}.json').read().strip().splitlines()[-1])
    k=d['config']['kernel_ms']
    print('$name', 'value %.4g ms/step %.4f icp %.4f lcp %.4f'%(d['value'],d['ms_per_step'],d['roofline']['kernel_ms_per_launch'],k['lcp_score']['ms_per_step']))
except Exception as e:
    print('$name', 'no line', e); print(open('$OUT/${TAG}_${name}.err').read()[-400:])
PY
}
for wl in shard2k shard4k shard8k; do
  run ${wl}_default $wl HOP_X=0
  run ${wl}_v1 $wl HOP_FUSED_VARIANT=1
  run ${wl}_v2 $wl HOP_FUSED_VARIANT=2
  run ${wl}_v2_s1 $wl HOP_FUSED_VARIANT=2 HOP_FUSED_SLOTS=1
done
