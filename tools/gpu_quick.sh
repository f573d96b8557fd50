#!/bin/bash
# quick GPU visit: selected tests + bench lines (no profiler).  usage: bash tools/gpu_quick.sh <tag> "<pytest -k expr or empty>" workloads...
TAG="$1"; shift; KEXPR="$1"; shift; WLS="${@:-C2 headline}"
mkdir -p gpurun_out
if [ -n "$KEXPR" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q -k "$KEXPR" 2>&1 | grep -v "^sampled\|^Pair set\|^Congruent\|^Q size\|^num \|^object sym" | tail -15
fi
for WL in $WLS; do
  timeout 600 python bench.py --workload $WL --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_${WL}.json 2> gpurun_out/${TAG}_bench_${WL}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench_${WL}.json"))
    print("$WL", "value %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], "ms/step %.3f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"],
          {k: round(v["ms_per_step"], 3) for k, v in d["config"]["kernel_ms"].items()}, "iters %.2f" % d["config"]["mean_icp_iterations"])
except Exception as e:
    print("$WL bench failed", e); print(open("gpurun_out/${TAG}_bench_${WL}.err").read()[-2000:])
PY
done
