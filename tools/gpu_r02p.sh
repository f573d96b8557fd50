#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02p}
timeout 600 python -m pytest tests/test_gpu_lm.py tests/test_gpu_parity.py -m gpu -q -x > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/${TAG}_pytest.log
timeout 300 python tools/ilp_equal.py 2>&1 | tail -8
run() { local name=$1; shift; local wl=$1; shift
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
for ilp in 1 2 4; do
  run headline_ilp$ilp headline HOP_FUSED_ILP=$ilp
  run headline_v1_ilp$ilp headline HOP_FUSED_ILP=$ilp HOP_FUSED_VARIANT=1
  run C2_ilp$ilp C2 HOP_FUSED_ILP=$ilp
  run C2_s1_ilp$ilp C2 HOP_FUSED_ILP=$ilp HOP_FUSED_SLOTS=1
  run C2_v2_ilp$ilp C2 HOP_FUSED_ILP=$ilp HOP_FUSED_VARIANT=2
  run C4_ilp$ilp C4 HOP_FUSED_ILP=$ilp
done
run C2_s1_ilp4_prof C2 HOP_FUSED_ILP=4 HOP_FUSED_SLOTS=1 HOP_FUSED_PROFILE=1
run headline_ilp2_prof headline HOP_FUSED_ILP=2 HOP_FUSED_PROFILE=1
