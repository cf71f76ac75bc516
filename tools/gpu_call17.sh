#!/bin/bash
# view-pooling encoder: parity tests + device time at the reference's size
O=gpurun_out/c17; mkdir -p $O
timeout 600 python -m pytest tests/test_encoder_gpu.py -q -s --tb=short > $O/pytest_encoder.log 2>&1; grep -v "^$" $O/pytest_encoder.log | tail -60 | cut -c1-400
for agg in mlp_mean angle; do
  timeout 300 python tools/encoder_bench.py --aggregator $agg > $O/encoder_$agg.json 2> $O/encoder_$agg.err; cut -c1-1500 $O/encoder_$agg.json; tail -3 $O/encoder_$agg.err
done
timeout 300 python tools/encoder_bench.py --aggregator mlp_mean --chunk 8192 > $O/encoder_mlp_mean_chunk8192.json 2>> $O/encoder_mlp_mean.err; cut -c1-600 $O/encoder_mlp_mean_chunk8192.json
timeout 300 python tools/encoder_bench.py --aggregator mlp_mean --chunk 32768 > $O/encoder_mlp_mean_chunk32768.json 2>> $O/encoder_mlp_mean.err; cut -c1-600 $O/encoder_mlp_mean_chunk32768.json
