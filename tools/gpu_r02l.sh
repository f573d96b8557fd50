#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02l}
timeout 1500 python -m pytest tests -m gpu -q -s > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log
grep -E "passed|failed|FAILED" $OUT/${TAG}_pytest_gpu.log | tail -20
timeout 600 python tools/bench_stages.py --steps 10 --warmup 3 > $OUT/${TAG}_bench_stages_C2.json 2> $OUT/${TAG}_bench_stages_C2.err; echo "stages exit $?"
python - <<PY
import json
for l in open('$OUT/${TAG}_bench_stages_C2.json'):
    try: d=json.loads(l)
    except Exception: continue
    print(d['stage'][:70], '|', 'value %.4g %s'%(d['value'],d['unit']), '| e2e', d['e2e'].get('ms_per_call'), '| cfg', json.dumps(d['config'].get('stage_ms', ''))[:300])
PY
