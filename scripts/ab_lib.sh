#!/bin/bash
# usage: scripts/ab_lib.sh libA.so libB.so ...  -> short bench line per library build, interleaved twice
P=spatial-temporal-lidar-camera-calibration_b200
cp $P/libstlcalib.so /tmp/keep.so
for rep in 1 2; do
for l in "$@"; do
  echo "== $l"
  cp $l $P/libstlcalib.so
  python bench.py --no-cpu-baseline --no-extras --steps 30 --warmup 5 2>&1 | grep metric | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(round(d['value'],1), round(d['ms_per_step'],3), d['stage_ms_per_launch'])"
done
done
cp /tmp/keep.so $P/libstlcalib.so
