python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r02k_c2.json 2> gpurun_out/r02k_c2.err; tail -3 gpurun_out/r02k_c2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02k_c2.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['stage_ms_per_launch'], 'check', d.get('result_check'))
PY
STL_K1_CLK=1 python scripts/k1_clk.py 600 2>&1 | tail -2
