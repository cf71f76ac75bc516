#!/bin/bash
# Round-2 first GPU call: validate the opt-in pieces of round 1 and measure each of them.
O=gpurun_out/c1; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
(HOLO_RUN_UNVALIDATED=1 timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -30) > $O/pytest_all.log 2>&1
tail -5 $O/pytest_all.log
(HOLO_RUN_UNVALIDATED=1 timeout 300 python -m pytest tests/test_unet_gpu.py -q -s -k "fused_skip or split_kv" 2>&1 | grep -v "^$" | tail -30) > $O/pytest_unval.log 2>&1
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
run() { name=$1; shift; env "$@" $B > $O/bench_$name.json 2> $O/bench_$name.err; python - $O/bench_$name.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "value %.2f ms %.3f e2e %.2f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), {k["kernel"][:12]: round(k.get("ms_per_step",0),3) for k in (d.get("roofline_other_kernels") or {}).get("kernels",[])}, d["roofline"]["share_of_step_ms"])
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
tail -2 $O/bench_$name.err; }
run default X=1
run fuse HOLO_FUSE_SKIP=1
run kvsplit HOLO_ATTN_KV_SPLIT=auto
run pdl HOLO_PDL=1
run all HOLO_FUSE_SKIP=1 HOLO_ATTN_KV_SPLIT=auto HOLO_PDL=1
run fuse_kv HOLO_FUSE_SKIP=1 HOLO_ATTN_KV_SPLIT=auto
(HOLO_FUSE_SKIP=1 HOLO_ATTN_KV_SPLIT=auto timeout 600 python -m pytest tests/test_fullsize_gpu.py -q -s 2>&1 | grep -v "^$" | tail -15) > $O/pytest_full_fuse_kv.log 2>&1
tail -6 $O/pytest_full_fuse_kv.log
(HOLO_CONV_MAX_CHAIN=54 timeout 600 python -m pytest tests/test_fullsize_gpu.py -q -s 2>&1 | grep -v "^$" | tail -15) > $O/pytest_full_chain54.log 2>&1
tail -6 $O/pytest_full_chain54.log
run chain54 HOLO_CONV_MAX_CHAIN=54
