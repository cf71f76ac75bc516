#!/bin/bash
# view-pooling encoder after the float4 column-group rewrite of the gather kernels
O=gpurun_out/c18; mkdir -p $O
timeout 600 python -m pytest tests/test_encoder_gpu.py -q -s --tb=short > $O/pytest_encoder.log 2>&1; grep -v "^$" $O/pytest_encoder.log | tail -40 | cut -c1-400
for v in "mlp_mean 0" "angle 0" "mlp_mean 32768" "mlp_mean 65536"; do
  set -- $v
  timeout 300 python tools/encoder_bench.py --aggregator $1 --chunk $2 > $O/encoder_$1_$2.json 2> $O/encoder_$1_$2.err; cut -c1-1300 $O/encoder_$1_$2.json; tail -2 $O/encoder_$1_$2.err
done
timeout 300 python tests/diagnostics/encoder_eager_compare.py > $O/encoder_eager_compare.json 2> $O/encoder_eager_compare.err; cat $O/encoder_eager_compare.json; tail -3 $O/encoder_eager_compare.err
