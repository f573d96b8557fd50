#!/bin/bash
# where does a frame's time go?  stage bench (incl. the drop-in executable), then the executable under an ncu launch list
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02x}
export HOP_KEEP_FRAME_DIR=/tmp/hop_frame
timeout 900 python tools/bench_stages.py --steps 10 --warmup 3 > $OUT/${TAG}_bench_stages_C2.json 2> $OUT/${TAG}_bench_stages_C2.err; echo "stages exit $?"
python - <<PY
import json
for l in open('$OUT/${TAG}_bench_stages_C2.json'):
    try: d=json.loads(l)
    except Exception: continue
    print(d['stage'][:60], '|', 'value %.4g %s'%(d['value'],d['unit']), '| e2e', d['e2e'].get('ms_per_call'), '|', json.dumps(d['config'].get('stage_ms', ''))[:400])
PY
ls /tmp/hop_frame | head
BIN=icra20-hand-object-pose_b200/host/main_realdata_auto
$BIN /tmp/hop_frame/cfg.yaml 6 2>&1 | tail -30 > $OUT/${TAG}_main_out.txt; tail -25 $OUT/${TAG}_main_out.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches_main.csv $BIN /tmp/hop_frame/cfg.yaml 2 > $OUT/${TAG}_ncu_main.log 2>&1; echo "ncu exit $?"
python tools/launch_summary.py $OUT/${TAG}_launches_main.csv > $OUT/${TAG}_launches_main.txt 2>&1; head -50 $OUT/${TAG}_launches_main.txt
