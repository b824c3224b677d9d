python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_c2.json 2> gpurun_out/r02_bench_n1_c2.err; tail -2 gpurun_out/r02_bench_n1_c2.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02_launches_bench.log 2>&1
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_n1_c2.json'))
print('value',round(d['value'],1),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value'],1),'cpu',d['cpu_baseline']['value'],'chk',d['oracle_check']['ok'],d['stage_ms_per_launch'], 'poll', d['poll_batch']['evals_per_s'], 'pf', d['plane_fit_per_query']['value'])
PY
