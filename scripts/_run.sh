python -m pytest tests/test_gpu_lm.py tests/test_gpu_parity.py -m gpu -x -q -s -k "gpr or golden or config5" 2>&1 | grep -v "^$" | tail -12
python scripts/sweep.py 6 12 2>&1 | tail -22
python bench.py --steps 10 --warmup 3 > gpurun_out/r02c_bench_c2.json 2> gpurun_out/r02c_bench_c2.err; tail -3 gpurun_out/r02c_bench_c2.err
python bench.py --config c4 --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/r02c_bench_c4.json 2> gpurun_out/r02c_bench_c4.err; tail -3 gpurun_out/r02c_bench_c4.err
python - <<'PY'
import json
for c in ('c2','c4'):
    d=json.load(open(f'gpurun_out/r02c_bench_{c}.json'))
    print(c, 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])
    print('   stages', d['stage_ms_per_launch'], 'roofline', d['roofline']['frac'], d['roofline']['avg_launch_ms'], d.get('poll_batch',{}).get('ms'), d.get('plane_fit_per_query',{}).get('ms_per_step'))
PY
