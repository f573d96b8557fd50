#!/bin/bash
# compute-sanitizer over the kernels round 2 added or rewrote: the warp-distributed LM (shared-memory hand-offs between lanes), the fused
# ICP with slots, the Kabsch path, the normals / MLS kernels, the device planner
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02y}
run() { local tool=$1 name=$2; shift 2
  timeout 600 compute-sanitizer --tool $tool --print-limit 5 python -m pytest "$@" -m gpu -q -x > $OUT/${TAG}_${tool}_${name}.txt 2>&1
  echo "$tool $name exit $? :" $(grep -c "Race reported\|hazard\|Invalid\|Error:" $OUT/${TAG}_${tool}_${name}.txt) "reports"; tail -3 $OUT/${TAG}_${tool}_${name}.txt
}
run racecheck lm tests/test_gpu_lm.py -k "cuboid or ellipse"
run racecheck icp tests/test_gpu_parity.py -k "in_the_convergence_basin and cuboid or runaway or pipelines_agree and ellipse"
run racecheck p2p tests/test_gpu_p2p.py -k "semantics"
run memcheck lm_icp tests/test_gpu_lm.py tests/test_gpu_parity.py -k "cuboid and (follows or convergence_basin) or runaway or golden"
run memcheck p2p_normals tests/test_gpu_p2p.py tests/test_gpu_normals.py -k "semantics or organized or mls"
run memcheck plan tests/test_gpu_s4pcs.py -k "device_assisted_plan or device_ppf_table"
run initcheck lm tests/test_gpu_lm.py -k "cuboid"
