#!/bin/bash
# round 2 visit h (1 GPU): full parity suite, smoke, default bench line (headline + C2 + stages), reference arm, ncu launch list + full captures
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02u}
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log
grep -E "passed|failed|FAILED" $OUT/${TAG}_pytest_gpu.log | tail -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?" >> $OUT/${TAG}_smoke.log; tail -2 $OUT/${TAG}_smoke.log
timeout 1200 python bench.py > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err; echo "bench default exit $?"; tail -c 1500 $OUT/${TAG}_bench_default.err
python - <<PY
import json
d=json.loads(open('$OUT/${TAG}_bench_default.json').read().strip().splitlines()[-1])
print('headline value %.4g ms/step %.3f e2e %.4g frac %.3f cpu %s'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac'],d.get('cpu_baseline')))
a=d['config']['also']['C2']; print('C2 value %.4g ms/step %.3f e2e %.4g frac %.3f'%(a['value'],a['ms_per_step'],a['e2e']['value'],a['roofline']['frac']))
print(json.dumps(d['config']['stage_ms'].get('frame_stages_C2_sizes'))[:1500])
PY
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err; echo "reference arm exit $?"; head -c 600 $OUT/${TAG}_bench_reference.json
for WL in headline C2; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/${TAG}_launches_${WL}.csv \
    python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline --no-also --no-stages > $OUT/${TAG}_ncu_launch_${WL}.log 2>&1
  echo "ncu launches $WL exit $?"
  for K in icp_fused_kernel lcp_score_kernel; do
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o $OUT/${TAG}_${K}_${WL} \
      python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline --no-also --no-stages > $OUT/${TAG}_ncu_${K}_${WL}.log 2>&1
    echo "ncu $K $WL exit $?"
    python tools/ncu_summary.py $OUT/${TAG}_${K}_${WL}.ncu-rep > $OUT/${TAG}_ncu_${K}_${WL}.txt 2>&1
  done
done
ls -la $OUT | grep $TAG | head -40
