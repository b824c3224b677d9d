python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/r02e_bench_c2.json 2> gpurun_out/r02e_bench_c2.err; tail -3 gpurun_out/r02e_bench_c2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02e_bench_c2.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['stage_ms_per_launch'], d['oracle_check']['ok'], d['oracle_check']['rel_err'])
PY
