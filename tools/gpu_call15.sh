#!/bin/bash
O=gpurun_out/c15; mkdir -p $O
(timeout 300 python -m pytest tests/test_model_gpu.py tests/test_unet_gpu.py -q -x -k "view_stream or splitk or native" 2>&1 | tail -3) > $O/pytest.log 2>&1; cat $O/pytest.log
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline"
show() { python - $1 $2 <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "value %.2f ms %.3f e2e %.2f (%.3f ms)" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]), "conv %.3f" % d["roofline"]["ms_per_step"], "sum %.2f" % d["roofline_other_kernels"]["sum_of_instrumented_kernels_ms_per_step"])
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
run() { name=$1; shift; env "$@" $B > $O/bench_$name.json 2> $O/bench_$name.err; show $O/bench_$name.json $name; tail -2 $O/bench_$name.err; }
run default X=1
run target444 HOLO_SPLITK_TARGET=444
run target592 HOLO_SPLITK_TARGET=592
run target148 HOLO_SPLITK_TARGET=148
run default_again X=1
