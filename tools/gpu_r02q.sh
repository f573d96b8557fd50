#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02q}; N=${2:-2}
nvidia-smi -L > $OUT/${TAG}_gpus.txt
timeout 300 python tools/lm_latency.py 2>&1 | tail -8
timeout 600 python -m pytest tests/test_gpu_lm.py tests/test_gpu_p2p.py tests/test_gpu_multi.py -m gpu -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/${TAG}_pytest.log
for SC in weak strong; do
  for WL in headline C2; do
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $WL --scaling $SC --steps 10 --warmup 3 --no-also --no-stages --no-cpu-baseline > $OUT/${TAG}_bench_${WL}_${SC}_${N}gpu.json 2> $OUT/${TAG}_bench_${WL}_${SC}_${N}gpu.err
    python - <<PY
import json
try:
    d=json.loads(open('$OUT/${TAG}_bench_${WL}_${SC}_${N}gpu.json').read().strip().splitlines()[-1])
    print('$WL $SC N=$N value %.4g ms/step %.4f e2e %.4g'%(d['value'],d['ms_per_step'],d['e2e']['value']), d['config'].get('collective'))
except Exception as e:
    print('no line', e); print(open('$OUT/${TAG}_bench_${WL}_${SC}_${N}gpu.err').read()[-1500:])
PY
  done
done
