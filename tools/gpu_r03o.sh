#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r03o}
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/${TAG}_pytest.log
for wl in headline C2 C4 C3; do python bench.py --workload $wl --steps 8 --warmup 3 --no-also --no-stages --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['config']['kernel_ms']; print(d['config']['workload'], 'value %.4g ms/step %.4f e2e %.4g icp %.4f lcp %.4f nn_build %.4f'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['kernel_ms_per_launch'],k['lcp_score']['ms_per_step'],k['nn_build']['ms_per_step']))"; done
export HOP_KEEP_FRAME_DIR=/tmp/hop_frame
timeout 900 python tools/bench_stages.py --steps 3 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_bench_stages.json 2> $OUT/${TAG}_bench_stages.err
icra20-hand-object-pose_b200/host/main_realdata_auto /tmp/hop_frame/cfg.yaml 20 2>/dev/null | grep timing_ms | tail -1
