#!/bin/bash
# Round-2 measurement campaign on one box with N GPUs visible: bash scripts/measure_r02.sh N
# N = 1: every BASELINE config (bench lines), the ncu launch list and full captures of the top kernels.
# N > 1: the default bench (configs[1], keyframes sharded, record exchanged by the library) and the poll / k = 20 configs.
N=${1:-1}
O=gpurun_out
if [ "$N" = "1" ]; then
  for c in c2 c1 c3 c4 c5; do
    st=20; [ $c = c4 ] && st=5; [ $c = c5 ] && st=8
    python bench.py --config $c --steps $st --warmup 5 > $O/r02_bench_n1_$c.json 2> $O/r02_bench_n1_$c.err
  done
  python bench.py --impl reference --steps 5 --warmup 1 > $O/r02_bench_reference_c2.json 2> $O/r02_bench_reference_c2.err
  ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $O/r02_launches_bench.log 2>&1
  for spec in "k1:k_assoc2d:2" "k2:k_nn_knn:2" "lm:k_lm_plane_b:2" "lin:k_linearize:2" "k0:k_kd_refine:1" "kidx:k_index_knn:3"; do
    IFS=: read name pat skip <<< "$spec"
    ncu --set full --clock-control none --import-source on -k regex:$pat -s $skip -c 1 -o $O/r02_${name}_full -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $O/r02_ncu_$name.log 2>&1
  done
  ncu -i $O/r02_k1_full.ncu-rep --page source --csv --print-source sass > $O/r02_k1_src.csv 2>/dev/null
  ncu --set full --clock-control none -k regex:k_assoc2d -s 6 -c 1 -o $O/r02_k1poll_full -f python bench.py --config c4 --steps 1 --warmup 1 --no-cpu-baseline > $O/r02_ncu_k1poll.log 2>&1
else
  for c in c2 c4 c3; do
    st=20; [ $c = c4 ] && st=5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --config $c --steps $st --warmup 5 2> $O/r02_bench_n${N}_$c.err | grep '^{' > $O/r02_bench_n${N}_$c.json
  done
  STL_NO_P2P=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 20 --warmup 5 --no-extras 2> /dev/null | grep '^{' > $O/r02_bench_n${N}_c2_nccl.json
  NCCL_DEBUG=INFO python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus $N --steps 3 --warmup 3 --no-extras 2>&1 | grep -E "NCCL INFO.*(Init COMPLETE|NVLS|nranks)" | head -6 > $O/r02_nccl_n$N.log
fi
python - <<PY
import json, glob
for f in sorted(glob.glob('$O/r02_bench_n${N}_*.json')):
    try:
        d = json.load(open(f))
        print(f.split('/')[-1], 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'], 1),
              'cpu', d.get('cpu_baseline', {}).get('value'), 'check', d.get('oracle_check', {}).get('ok'), d['stage_ms_per_launch'])
    except Exception as e:
        print(f, 'FAILED', e)
PY
