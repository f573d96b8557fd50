#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r03e}
timeout 900 python -m pytest tests/test_gpu_s4pcs.py tests/test_gpu_fullsize.py tests/test_gpu_pipeline_physics.py tests/test_gpu_main_hand.py -m gpu -q -x > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/${TAG}_pytest.log
export HOP_KEEP_FRAME_DIR=/tmp/hop_frame
timeout 900 python tools/bench_stages.py --steps 3 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_bench_stages.json 2> $OUT/${TAG}_bench_stages.err; echo "stages exit $?"
BIN=icra20-hand-object-pose_b200/host/main_realdata_auto
HOP_TRACE=1 $BIN /tmp/hop_frame/cfg.yaml 20 > $OUT/${TAG}_main_out.txt 2> $OUT/${TAG}_main_trace.txt; grep timing_ms $OUT/${TAG}_main_out.txt | tail -3; grep "hop trace" $OUT/${TAG}_main_trace.txt | head -16
