#!/bin/bash
# the bench line with the view-pooling encoder evidence leg
O=gpurun_out/c23; mkdir -p $O
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err; tail -2 $O/bench.err
python - $O/bench.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"])
print("encoder", d["view_pooling_encoder"])
print("eager", {k:v for k,v in (d["gpu_eager_baseline"] or {}).items() if k!="what"})
PY
