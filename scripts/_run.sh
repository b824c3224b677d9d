python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --config c5 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r02y_c5.json 2> gpurun_out/r02y_c5.err; tail -3 gpurun_out/r02y_c5.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02y_c5.json'))
print('c5 value',round(d['value'],1),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value'],1),d['stage_ms_per_launch'], d['result_check'])
PY
