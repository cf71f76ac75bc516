#!/bin/bash
# Round-2 call 2: chunked TMEM accumulation in conv_tc -- correctness, error vs chunk length, cost.
O=gpurun_out/c2; mkdir -p $O
(timeout 600 python -m pytest tests/test_unet_gpu.py -q -x 2>&1 | tail -8) > $O/pytest_unet.log 2>&1
tail -3 $O/pytest_unet.log
for c in 0 1 2 4 8; do
  HOLO_CONV_CHUNK=$c timeout 300 python tests/diagnostics/unet_error_trace.py --f64 > $O/trace_chunk$c.log 2>&1
  echo "chunk $c: $(tail -2 $O/trace_chunk$c.log | tr '\n' ' ')"
done
HOLO_CONV_CHUNK=4 HOLO_ATTN_KV_SPLIT=8 timeout 300 python tests/diagnostics/unet_error_trace.py --f64 > $O/trace_chunk4_kv8.log 2>&1
echo "chunk 4 kv8: $(tail -2 $O/trace_chunk4_kv8.log | tr '\n' ' ')"
timeout 300 python tests/diagnostics/unet_error_trace.py --f64 --no-tc > $O/trace_notc.log 2>&1
echo "no-tc: $(tail -2 $O/trace_notc.log | tr '\n' ' ')"
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
for c in 0 2 4 8; do
  HOLO_CONV_CHUNK=$c $B > $O/bench_chunk$c.json 2> $O/bench_chunk$c.err
  python - $O/bench_chunk$c.json chunk$c <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "value %.2f ms %.3f e2e %.2f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["roofline"]["share_of_step_ms"])
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
done
(timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5) > $O/pytest_all.log 2>&1
tail -3 $O/pytest_all.log
