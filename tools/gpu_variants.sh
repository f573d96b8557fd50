#!/bin/bash
# fused-ICP CTA-shape sweep: bash tools/gpu_variants.sh <tag> <workload> variants...
TAG="$1"; shift; WL="$1"; shift
mkdir -p gpurun_out
for V in "$@"; do
  HOP_FUSED_VARIANT=$V timeout 300 python bench.py --workload $WL --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_v${V}_${WL}.json 2> gpurun_out/${TAG}_v${V}_${WL}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_v${V}_${WL}.json"))
    print("variant $V $WL", "value %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], "ms/step %.3f" % d["ms_per_step"], {k: round(v["ms_per_step"], 4) for k, v in d["config"]["kernel_ms"].items()})
except Exception as e:
    print("variant $V failed", e); print(open("gpurun_out/${TAG}_v${V}_${WL}.err").read()[-1500:])
PY
done
