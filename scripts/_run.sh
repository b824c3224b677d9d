python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for b in 4 8 16; do
STL_K2_BATCH=$b python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r02s_c2_b$b.json 2> gpurun_out/r02s_c2_b$b.err
done
for b in 2 4 8; do
STL_K2_BATCH=$b python bench.py --nkf 188 --steps 40 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r02s_188_b$b.json 2> gpurun_out/r02s_188_b$b.err
done
python - <<'PY'
import json
for n in ('c2_b4','c2_b8','c2_b16','188_b2','188_b4','188_b8'):
    d=json.load(open(f'gpurun_out/r02s_{n}.json'))
    print(n,'value',round(d['value'],1),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value'],1),d['stage_ms_per_launch'])
PY
