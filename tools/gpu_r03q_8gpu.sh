#!/bin/bash
# 8-GPU box: scaling lines through the C-ABI collective (weak = own frames per rank, strong = one frame's hypotheses sharded)
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02w}
nvidia-smi -L > $OUT/${TAG}_gpus.txt
line() { # N workload scaling
  local N=$1 WL=$2 SC=$3
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $WL --scaling $SC --steps 10 --warmup 3 --no-also --no-stages --no-cpu-baseline > $OUT/${TAG}_bench_${WL}_${SC}_${N}gpu.json 2> $OUT/${TAG}_bench_${WL}_${SC}_${N}gpu.err
  python - <<PY
import json
try:
    d=json.loads(open('$OUT/${TAG}_bench_${WL}_${SC}_${N}gpu.json').read().strip().splitlines()[-1])
    print('$WL $SC N=$N value %.4g ms/step %.4f e2e %.4g'%(d['value'],d['ms_per_step'],d['e2e']['value']), {k:v for k,v in (d['config'].get('collective') or {}).items() if k!='what'})
except Exception as e:
    print('$WL $SC N=$N no line', e); print(open('$OUT/${TAG}_bench_${WL}_${SC}_${N}gpu.err').read()[-800:])
PY
}
line 8 headline weak
line 8 headline strong
line 4 headline weak
line 4 headline strong
line 8 C2 weak
