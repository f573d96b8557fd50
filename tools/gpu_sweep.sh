#!/bin/bash
# sweep of a tuning knob (environment variable) over bench workloads.  usage: bash tools/gpu_sweep.sh <tag> <ENVVAR> "<values>" workloads...
TAG="$1"; shift; VAR="$1"; shift; VALS="$1"; shift; WLS="${@:-C2 headline}"
mkdir -p gpurun_out
for V in $VALS; do
  for WL in $WLS; do
    env $VAR=$V timeout 600 python bench.py --workload $WL --steps 10 --warmup 4 --no-cpu-baseline > gpurun_out/${TAG}_${V}_${WL}.json 2> gpurun_out/${TAG}_${V}_${WL}.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_${V}_${WL}.json"))
    print("$VAR=$V $WL", "value %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], "ms/step %.3f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"],
          {k: round(v["ms_per_step"], 3) for k, v in d["config"]["kernel_ms"].items()}, "iters %.2f" % d["config"]["mean_icp_iterations"])
except Exception as e:
    print("$VAR=$V $WL bench failed", e); print(open("gpurun_out/${TAG}_${V}_${WL}.err").read()[-1500:])
PY
  done
done
