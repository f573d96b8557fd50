#!/bin/bash
# small-batch experiments: CTA shape / slots per CTA at C2 and C4 (one frame = 4096 hypotheses)
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02n}
run() { # name, env..., workload
  local name=$1; shift; local wl=$1; shift
  env "$@" timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-also --no-stages --no-cpu-baseline > $OUT/${TAG}_${name}.json 2> $OUT/${TAG}_${name}.err
  python - <<PY
import json
try:
    d=json.loads(open('$OUT/${TAG}_${name}.json').read().strip().splitlines()[-1])
    print('$name', 'value %.4g ms/step %.4f icp %.4f'%(d['value'],d['ms_per_step'],d['roofline']['kernel_ms_per_launch']))
except Exception as e:
    print('$name', 'no line', e); print(open('$OUT/${TAG}_${name}.err').read()[-600:])
PY
  grep "hop fused profile" $OUT/${TAG}_${name}.err | tail -1
}
run C2_default C2 HOP_X=0
run C2_v2 C2 HOP_FUSED_VARIANT=2
run C2_v2_prof C2 HOP_FUSED_VARIANT=2 HOP_FUSED_PROFILE=1
run C2_v1_s1 C2 HOP_FUSED_VARIANT=1 HOP_FUSED_SLOTS=1
run C4_default C4 HOP_X=0
run C4_v2_s1 C4 HOP_FUSED_VARIANT=2 HOP_FUSED_SLOTS=1
run C4_v2_s2 C4 HOP_FUSED_VARIANT=2 HOP_FUSED_SLOTS=2
run C4_v1 C4 HOP_FUSED_VARIANT=1
run C4_v1_s1 C4 HOP_FUSED_VARIANT=1 HOP_FUSED_SLOTS=1
run C3_v1 C3 HOP_FUSED_VARIANT=1
run C3_v2_s1 C3 HOP_FUSED_VARIANT=2 HOP_FUSED_SLOTS=1
run C3_v2_s2 C3 HOP_FUSED_VARIANT=2 HOP_FUSED_SLOTS=2
