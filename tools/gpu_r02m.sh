#!/bin/bash
# full -m gpu suite, smoke, default bench, the other BASELINE workloads with the parity solver
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02m}
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > $OUT/${TAG}_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log
grep -E "passed|failed|FAILED|pytest exit" $OUT/${TAG}_pytest_gpu.log | tail -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 $OUT/${TAG}_smoke.log
timeout 900 python bench.py > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err; echo "bench default exit $?"
for WL in C3 C4 C5; do
  timeout 900 python bench.py --workload $WL --steps 5 --warmup 3 --no-also --no-stages > $OUT/${TAG}_bench_${WL}.json 2> $OUT/${TAG}_bench_${WL}.err; echo "bench $WL exit $?"
done
python - <<PY
import json
for n in ['default','C3','C4','C5']:
    try:
        d=json.loads(open('$OUT/${TAG}_bench_%s.json'%n).read().strip().splitlines()[-1])
        print(n,'value %.4g ms/step %.4f e2e %.4g frac %.3f'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac']), 'cpu', d.get('cpu_baseline',{}).get('value'))
    except Exception as e:
        print(n,'no line',e); print(open('$OUT/${TAG}_bench_%s.err'%n).read()[-800:])
PY
