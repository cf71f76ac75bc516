#!/bin/bash
O=gpurun_out/c7; mkdir -p $O
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
run chunk0 HOLO_CONV_CHUNK=0
run default_again X=1
(timeout 600 python -m pytest tests/test_unet_gpu.py tests/test_model_gpu.py tests/test_postprocess_gpu.py tests/test_chain_drift_gpu.py -q -x -s 2>&1 | grep -E "passed|failed|Error|final_rel" | tail -6) > $O/pytest.log 2>&1; cat $O/pytest.log
timeout 300 python tools/one_sample_multi_gpu.py --resol 32 --image 128 --pts 64 > $O/cfg5_32_1gpu.json 2> $O/cfg5_32.err; cat $O/cfg5_32_1gpu.json; tail -2 $O/cfg5_32.err
timeout 600 python tools/one_sample_multi_gpu.py --resol 64 --image 256 --pts 64 > $O/cfg5_64_1gpu.json 2> $O/cfg5_64.err; cat $O/cfg5_64_1gpu.json; tail -2 $O/cfg5_64.err
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
