#!/bin/bash
# Round-2 call 8 (1 GPU): native whole-graph UNet (holo_unet_fwd), full suite, host-overhead A/B, cfg#4 1-GPU reference.
O=gpurun_out/c8; mkdir -p $O
(timeout 600 python -m pytest tests/test_unet_gpu.py -q -x -s -k "native" 2>&1 | grep -E "native UNet|passed|failed|Error|error" | tail -12) > $O/pytest_native.log 2>&1; cat $O/pytest_native.log
(timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6) > $O/pytest_all.log 2>&1; tail -3 $O/pytest_all.log
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline"
show() { python - $1 $2 <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "value %.2f ms %.3f e2e %.2f (%.3f ms)" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
run() { name=$1; shift; env "$@" $B "${EXTRA[@]}" > $O/bench_$name.json 2> $O/bench_$name.err; show $O/bench_$name.json $name; tail -2 $O/bench_$name.err; }
EXTRA=(); run python_graph X=1
EXTRA=(); run native_graph HOLO_UNET_NATIVE=1
EXTRA=(--no-graph); run python_eager X=1
EXTRA=(--no-graph); run native_eager HOLO_UNET_NATIVE=1
timeout 300 python tools/batch_sharded.py --batch 32 --repeats 2 > $O/cfg4_1gpu.json 2> $O/cfg4_1gpu.err; cat $O/cfg4_1gpu.json; tail -2 $O/cfg4_1gpu.err
timeout 300 python tools/sample_turntable.py > $O/cfg3.json 2> $O/cfg3.err; cat $O/cfg3.json; tail -2 $O/cfg3.err
HOLO_UNET_NATIVE=1 timeout 300 python tools/sample_turntable.py > $O/cfg3_native.json 2> $O/cfg3_native.err; cat $O/cfg3_native.json; tail -2 $O/cfg3_native.err
