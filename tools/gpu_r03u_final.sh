#!/bin/bash
# last visit: the whole -m gpu suite, smoke, and memcheck / racecheck over what changed after the first sanitizer pass
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r03u}
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -2 $OUT/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -1 $OUT/${TAG}_smoke.log
run() { local tool=$1 name=$2; shift 2
  timeout 600 compute-sanitizer --tool $tool --print-limit 5 python -m pytest "$@" -m gpu -q -x > $OUT/${TAG}_${tool}_${name}.txt 2>&1
  echo "$tool $name exit $?"; tail -2 $OUT/${TAG}_${tool}_${name}.txt
}
run memcheck cluster_plan tests/test_gpu_fullsize.py tests/test_gpu_s4pcs.py -k "cluster_poses or device_assisted_plan or device_ppf_table"
run racecheck lm tests/test_gpu_lm.py
run memcheck lm_static tests/test_gpu_lm.py tests/test_gpu_parity.py -k "warp_lm or static_hint or runaway"
