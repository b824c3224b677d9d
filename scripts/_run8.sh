N=8; O=gpurun_out
for c in c2 c4; do
  st=20; [ $c = c4 ] && st=5
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --config $c --steps $st --warmup 5 --no-cpu-baseline 2> $O/r02_bench_n${N}_$c.err | grep '^{' > $O/r02_bench_n${N}_$c.json
done
STL_NO_P2P=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 20 --warmup 5 --no-extras --no-cpu-baseline 2> /dev/null | grep '^{' > $O/r02_bench_n${N}_c2_nccl.json
NCCL_DEBUG=INFO STL_NO_P2P=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus $N --steps 3 --warmup 3 --no-extras --no-cpu-baseline 2>&1 | grep -E "NCCL INFO.*(Init COMPLETE|NVLS|nranks|Connected)" | head -8 > $O/r02_nccl_n$N.log
python - <<PY
import json, glob
for f in sorted(glob.glob('$O/r02_bench_n${N}_*.json')):
    try:
        d = json.load(open(f))
        print(f.split('/')[-1], 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'], 1), d['stage_ms_per_launch'], (d.get('poll_batch') or {}).get('evals_per_s'))
    except Exception as e:
        print(f, 'FAILED', e)
PY
