#!/bin/bash
# Round-2 call 10: chunked O accumulation in the fused attention, renderer backward kernels, full suite, bench.
O=gpurun_out/c10; mkdir -p $O
(timeout 500 python -m pytest tests/test_unet_gpu.py -q -x -s -k "chunked_o or attention_flash" 2>&1 | grep -vE "^$" | tail -12) > $O/pytest_attn.log 2>&1; cat $O/pytest_attn.log
(timeout 400 python -m pytest tests/test_render_bwd_gpu.py -q -s 2>&1 | grep -E "backward|passed|failed|Error|assert" | tail -14) > $O/pytest_bwd.log 2>&1; cat $O/pytest_bwd.log
(timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6) > $O/pytest_all.log 2>&1; tail -3 $O/pytest_all.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > $O/bench.json 2> $O/bench.err; cut -c1-300 $O/bench.json; tail -2 $O/bench.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/c10/bench.json").read().strip().splitlines()[-1])
    print("conv %.3f" % d["roofline"]["ms_per_step"], {k["kernel"][:10]: round(k["ms_per_step"],3) for k in d["roofline_other_kernels"]["kernels"]}, "sum %.2f" % d["roofline_other_kernels"]["sum_of_instrumented_kernels_ms_per_step"], "launches", d["gpu_launches"])
except Exception as e: print("FAILED", e)
PY
