#!/bin/bash
# One GPU-box visit: parity tests, bench lines, ncu launch list + one full capture of the dominant kernels.
# usage (under gpurun): bash tools/gpu_round.sh <tag> [tests|notests] [workloads...]
TAG="${1:-r01}"; shift
TESTS="${1:-tests}"; shift
WLS="${@:-C2}"
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt 2>&1
if [ "$TESTS" = "tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1
  echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log
  tail -5 $OUT/${TAG}_pytest_gpu.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
  echo "smoke exit $?" >> $OUT/${TAG}_smoke.log
  tail -3 $OUT/${TAG}_smoke.log
fi
for WL in $WLS; do
  timeout 900 python bench.py --workload $WL --steps 10 --warmup 3 > $OUT/${TAG}_bench_${WL}.json 2> $OUT/${TAG}_bench_${WL}.err
  echo "bench $WL exit $?"; tail -c 3000 $OUT/${TAG}_bench_${WL}.json; tail -3 $OUT/${TAG}_bench_${WL}.err
done
WL0=$(echo $WLS | cut -d' ' -f1)
# launch list of the same command (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches_${WL0}.csv \
  python bench.py --workload $WL0 --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_launch.log 2>&1
echo "ncu launches exit $?"
# full capture of the hot kernels (3 launches each, after warm-up)
for K in icp_fused_kernel lcp_score_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 2 -f -o $OUT/${TAG}_${K}_${WL0} \
    python bench.py --workload $WL0 --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_${K}.log 2>&1
  echo "ncu $K exit $?"
done
ls -la $OUT
