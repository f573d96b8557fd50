#!/bin/bash
# N GPUs: the two-contexts and two-ranks tests, then weak / strong scaling bench lines through the C ABI collective
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02i}; N=${2:-2}
nvidia-smi -L > $OUT/${TAG}_gpus.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -s > $OUT/${TAG}_pytest_multi.log 2>&1; echo "pytest multi exit $?"; tail -5 $OUT/${TAG}_pytest_multi.log
for SC in weak strong; do
  for WL in headline C2; do
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $WL --scaling $SC --steps 10 --warmup 3 --no-also --no-stages > $OUT/${TAG}_bench_${WL}_${SC}_${N}gpu.json 2> $OUT/${TAG}_bench_${WL}_${SC}_${N}gpu.err
    echo "bench $WL $SC N=$N exit $?"
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
