python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r03d_c2.json 2> gpurun_out/r03d_c2.err; tail -3 gpurun_out/r03d_c2.err
python bench.py --nkf 188 --steps 40 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r03d_c2_188.json 2> gpurun_out/r03d_c2_188.err
python bench.py --config c4 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r03d_c4.json 2> gpurun_out/r03d_c4.err
python - <<'PY'
import json
for n in ('r03d_c2','r03d_c2_188','r03d_c4'):
    d=json.load(open(f'gpurun_out/{n}.json'))
    print(n,'value',round(d['value'],1),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value'],1),d['stage_ms_per_launch'])
PY
