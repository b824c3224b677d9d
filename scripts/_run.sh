timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -8
for e in "STL_X=1" "STL_NO_P2P=1"; do
env $e timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --no-extras 2>/dev/null | grep '^{' | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('$e', round(d['value'],1), round(d['ms_per_step'],4), d['stage_ms_per_launch'])"
done
