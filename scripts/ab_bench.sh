#!/bin/bash
# usage: scripts/ab_bench.sh "ENV1=.. ENV2=.." "ENV.." ...   -> one short bench line per environment, interleaved twice
for rep in 1 2; do
for e in "$@"; do
  echo "== $e"
  env $e python bench.py --no-cpu-baseline --no-extras --steps 30 --warmup 5 2>&1 | grep metric | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(round(d['value'],1), round(d['ms_per_step'],3), d['stage_ms_per_launch'])"
done
done
