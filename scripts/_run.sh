python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for e in "STL_SUB=4" "STL_SUB=8" "STL_SUB=16"; do env $e python bench.py --config c2 --nkf 188 --steps 60 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$e', round(d['value'],1), round(d['ms_per_step'],4), d['stage_ms_per_launch'])"; done
