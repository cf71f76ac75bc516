#!/bin/bash
# Round-2 call 3: cost of chunked accumulation with batched TMEM loads (LDW 16 / 32) at chunk 4 / 8.
O=gpurun_out/c3; mkdir -p $O
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
show() { python - $1 $2 <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "value %.2f ms %.3f e2e %.2f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["roofline"]["share_of_step_ms"])
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
for w in 32 16; do
  HOLO_NVCC_FLAGS="-DHOLO_CONV_LDW=$w" python holo_diffusion_b200/build.py > $O/build_$w.log 2>&1 || tail -5 $O/build_$w.log
  for c in 0 4 8; do
    HOLO_CONV_CHUNK=$c $B > $O/bench_ldw${w}_chunk$c.json 2> $O/bench_ldw${w}_chunk$c.err
    show $O/bench_ldw${w}_chunk$c.json ldw${w}_chunk$c
  done
done
python holo_diffusion_b200/build.py > $O/build_default.log 2>&1
(timeout 600 python -m pytest tests/test_unet_gpu.py tests/test_fullsize_gpu.py -q -x -s 2>&1 | grep -E "tanh|passed|failed|Error" | tail -8) > $O/pytest.log 2>&1
cat $O/pytest.log
HOLO_CONV_CHUNK=4 timeout 300 python tests/diagnostics/unet_error_trace.py --f64 2>&1 | tail -1
