#!/bin/bash
O=gpurun_out/c6; mkdir -p $O
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline"
show() { python - $1 $2 <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "value %.2f ms %.3f e2e %.2f (%.3f ms)" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]), "conv %.3f" % d["roofline"]["ms_per_step"], {k["kernel"][:10]: round(k["ms_per_step"],3) for k in d["roofline_other_kernels"]["kernels"]}, "sum %.2f" % d["roofline_other_kernels"]["sum_of_instrumented_kernels_ms_per_step"])
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
run() { name=$1; shift; env "$@" $B > $O/bench_$name.json 2> $O/bench_$name.err; show $O/bench_$name.json $name; tail -2 $O/bench_$name.err; }
run default X=1
run nostats HOLO_SPLITK_STATS=0
run noclone HOLO_GRAPH_NO_CLONE=1
run nostats_noclone HOLO_SPLITK_STATS=0 HOLO_GRAPH_NO_CLONE=1
run nostats_noclone_chunk0 HOLO_SPLITK_STATS=0 HOLO_GRAPH_NO_CLONE=1 HOLO_CONV_CHUNK=0
