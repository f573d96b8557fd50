#!/bin/bash
# GPU visit for the other kernels of the path: stage benches (K1, Super4PCS) + ncu launch list and full captures.
TAG="${1:-r01}"; OUT=gpurun_out; mkdir -p $OUT
for SZ in C2 C5; do
  timeout 900 python tools/bench_stages.py --sizes $SZ > $OUT/${TAG}_stages_${SZ}.json 2> $OUT/${TAG}_stages_${SZ}.err
  echo "stages $SZ exit $?"; cut -c1-700 $OUT/${TAG}_stages_${SZ}.json; tail -2 $OUT/${TAG}_stages_${SZ}.err
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_stages_C2.csv \
  python tools/bench_stages.py --sizes C2 --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_stages_launch.log 2>&1
echo "ncu launches exit $?"
for K in hand_overlap_kernel verify_lcp_kernel extract_pairs_kernel congruent_join_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o $OUT/${TAG}_${K}_C2 \
    python tools/bench_stages.py --sizes C2 --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_${K}.log 2>&1
  echo "ncu $K exit $?"
done
