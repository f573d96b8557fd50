#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02j}
timeout 1500 python -m pytest tests -m gpu -q -s > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log
grep -E "passed|failed|FAILED" $OUT/${TAG}_pytest_gpu.log | tail -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?" >> $OUT/${TAG}_smoke.log; tail -2 $OUT/${TAG}_smoke.log
for WL in C2 headline; do
  timeout 900 python bench.py --workload $WL --steps 10 --warmup 3 --no-cpu-baseline --no-also --no-stages > $OUT/${TAG}_bench_${WL}.json 2> $OUT/${TAG}_bench_${WL}.err
  python - <<PY
import json
d=json.loads(open('$OUT/${TAG}_bench_${WL}.json').read().strip().splitlines()[-1])
k=d['config']['kernel_ms']
print('$WL', 'value %.4g ms/step %.4f e2e %.4g frac %.3f'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac']), {a:round(b['ms_per_step'],4) for a,b in k.items()})
PY
done
