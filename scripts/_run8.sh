N=$1
for c in c2 c3 c4; do
  st=20; [ $c = c4 ] && st=5
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --config $c --steps $st --warmup 5 2> gpurun_out/r02f_bench_n${N}_$c.err | grep '^{' > gpurun_out/r02f_bench_n${N}_$c.json
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02f_bench_n${N}_$c.json'))
    print('$c N=$N value', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), d['stage_ms_per_launch'], d.get('poll_batch',{}).get('ms'))
except Exception as e:
    print('$c FAILED', e)
PY
done
NCCL_DEBUG=INFO python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 5 --warmup 3 --no-extras 2>&1 | grep -E "NCCL INFO (comm|ncclCommInitRank|Connected|NVLS|.*AllReduce)" | grep -v "torch" | head -8 > gpurun_out/r02f_nccl_n$N.log
tail -3 gpurun_out/r02f_nccl_n$N.log | cut -c1-200
