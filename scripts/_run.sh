python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py --steps 10 --warmup 3 > gpurun_out/r02b_bench_c2.json 2> gpurun_out/r02b_bench_c2.err; tail -3 gpurun_out/r02b_bench_c2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02b_bench_c2.json'))
for k in ('value','ms_per_step','e2e','gpu_launches','setup','stage_ms_per_launch','oracle_check','poll_batch','plane_fit_per_query'):
    print(k, d.get(k))
print(d['cpu_baseline']['value'], d['cpu_baseline'].get('bae_only'))
PY
