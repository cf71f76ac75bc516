#!/bin/bash
# encoder GEMMs: N = 64 tiles for the short-K Linear layers?
O=gpurun_out/c22; mkdir -p $O
for v in 0 1 0 1; do
  HOLO_GEMM_SHORTK_N64=$v timeout 300 python tools/encoder_bench.py --aggregator mlp_mean --iters 10 > $O/encoder_n64_$v.json 2> $O/encoder_$v.err; python - $O/encoder_n64_$v.json $v <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print("N64=%s"%sys.argv[2], d["ms_per_grid"], {k:v["ms"] for k,v in d["per_entry_point"].items()})
PY
done
HOLO_GEMM_SHORTK_N64=1 timeout 300 python -m pytest tests/test_encoder_gpu.py -q --tb=short -k "stages or full_size" 2>&1 | tail -2
