#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02z}
run() { local tool=$1 name=$2; shift 2
  timeout 600 compute-sanitizer --tool $tool --print-limit 5 python -m pytest "$@" -m gpu -q -x > $OUT/${TAG}_${tool}_${name}.txt 2>&1
  echo "$tool $name exit $?"; tail -2 $OUT/${TAG}_${tool}_${name}.txt
}
run racecheck icp tests/test_gpu_parity.py -k "in_the_convergence_basin and cuboid or runaway or pipelines_agree and ellipse or semantics_of_the_reference"
HOP_FUSED_SLOTS=3 run racecheck icp_slots3 tests/test_gpu_parity.py -k "in_the_convergence_basin and cuboid"
HOP_FUSED_VARIANT=2 HOP_FUSED_SLOTS=2 run racecheck icp_v2_slots2 tests/test_gpu_parity.py -k "in_the_convergence_basin and tless"
