#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r03l}
run() { local name=$1; shift; local wl=$1; shift
  env "$@" timeout 600 python bench.py --workload $wl --steps 6 --warmup 3 --no-also --no-stages --no-cpu-baseline > $OUT/${TAG}_${name}.json 2> $OUT/${TAG}_${name}.err
  python - <<PY
import json
try:
    d=json.loads(open('$OUT/${TAG}_${name}.json').read().strip().splitlines()[-1])
    k=d['config']['kernel_ms']
    print('$name', 'value %.4g ms/step %.4f icp %.4f lcp %.4f nn_build %.4f'%(d['value'],d['ms_per_step'],d['roofline']['kernel_ms_per_launch'],k['lcp_score']['ms_per_step'],k['nn_build']['ms_per_step']), d['config']['nn_grid_icp']['bytes'], d['config']['nn_grid_lcp']['bytes'])
except Exception as e:
    print('$name', 'no line', e); print(open('$OUT/${TAG}_${name}.err').read()[-400:])
PY
}
for mf in 1.0 0.7 0.5 0.35; do
  run headline_mf$mf headline HOP_VOXEL_MAX_FRAC=$mf
  run C4_mf$mf C4 HOP_VOXEL_MAX_FRAC=$mf
done
run C2_mf0.5 C2 HOP_VOXEL_MAX_FRAC=0.5
run C5_mf0.5 C5 HOP_VOXEL_MAX_FRAC=0.5
export HOP_KEEP_FRAME_DIR=/tmp/hop_frame
timeout 900 python tools/bench_stages.py --steps 3 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_bench_stages.json 2> $OUT/${TAG}_bench_stages.err
BIN=icra20-hand-object-pose_b200/host/main_realdata_auto
for mf in 1.0 0.5; do HOP_VOXEL_MAX_FRAC=$mf $BIN /tmp/hop_frame/cfg.yaml 20 2>/dev/null | grep timing_ms | tail -1; done
