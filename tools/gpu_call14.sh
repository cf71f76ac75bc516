#!/bin/bash
# Round-2 call 14: deterministic split-K as conv + reduce launch: correctness, A/B against the atomics path.
O=gpurun_out/c14; mkdir -p $O
(timeout 400 python -m pytest tests/test_unet_gpu.py -q -x -s -k "splitk or conv_tc or native or base_args" 2>&1 | grep -E "split-K|passed|failed|Error|assert" | tail -10) > $O/pytest_conv.log 2>&1; cat $O/pytest_conv.log
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline"
show() { python - $1 $2 <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "value %.2f ms %.3f e2e %.2f (%.3f ms)" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]), "conv %.3f" % d["roofline"]["ms_per_step"], {k["kernel"][:10]: round(k["ms_per_step"],3) for k in d["roofline_other_kernels"]["kernels"]}, "sum %.2f" % d["roofline_other_kernels"]["sum_of_instrumented_kernels_ms_per_step"], "launches", d["gpu_launches"])
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
run() { name=$1; shift; env "$@" $B > $O/bench_$name.json 2> $O/bench_$name.err; show $O/bench_$name.json $name; tail -2 $O/bench_$name.err; }
run atomics HOLO_SPLITK_WS=0
run ws HOLO_SPLITK_WS=1
run ws_max40 HOLO_SPLITK_WS=1 HOLO_SPLITK_MAX=40
run atomics_again HOLO_SPLITK_WS=0
run ws_again HOLO_SPLITK_WS=1
