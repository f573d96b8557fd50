#!/bin/bash
# round 2, visit a: parity suite with the LM-replay solver, bench C2 + headline, per-phase cycle accounting
OUT=gpurun_out; mkdir -p $OUT; TAG=r02a
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log
grep -E "passed|failed|FAILED|within bound|sample of" $OUT/${TAG}_pytest_gpu.log | tail -40
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?" >> $OUT/${TAG}_smoke.log; tail -3 $OUT/${TAG}_smoke.log
for WL in C2 headline; do
  for S in 0 2; do
    timeout 900 python bench.py --workload $WL --steps 10 --warmup 3 --solver $S --no-cpu-baseline > $OUT/${TAG}_bench_${WL}_s$S.json 2> $OUT/${TAG}_bench_${WL}_s$S.err
    echo "bench $WL solver $S exit $?"; python -c "
import json,sys
d=json.loads(open('$OUT/${TAG}_bench_${WL}_s$S.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','e2e','roofline','kernel_ms')})"
  done
  HOP_FUSED_PROFILE=1 timeout 600 python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | grep "fused profile" | tail -2
done
