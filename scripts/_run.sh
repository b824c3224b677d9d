python bench.py --config c4 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r02l_c4.json 2> gpurun_out/r02l_c4.err; tail -3 gpurun_out/r02l_c4.err
python bench.py --nkf 188 --steps 30 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r02l_c2_188.json 2> gpurun_out/r02l_c2_188.err
STL_K1_SPLIT=1 python bench.py --nkf 188 --steps 30 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r02l_c2_188_split.json 2> gpurun_out/r02l_c2_188_split.err
python - <<'PY'
import json
for n in ('r02l_c4','r02l_c2_188','r02l_c2_188_split'):
    d=json.load(open(f'gpurun_out/{n}.json'))
    print(n,'value',round(d['value'],1),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value'],1),d['stage_ms_per_launch'], 'k1 launches', d['roofline']['launches'], d['roofline'].get('candidates_per_launch'))
PY
