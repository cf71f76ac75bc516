R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline"
P='import json,sys
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"])'
echo noclocks; timeout 200 $R --no-clocks 2>/dev/null | python -c "$P"
echo e2efirst; timeout 200 $R --e2e-first 2>/dev/null | python -c "$P"
