#!/bin/bash
# encoder: whole grid as one chunk by default, launch attributes cached
O=gpurun_out/c21; mkdir -p $O
timeout 600 python -m pytest tests/test_encoder_gpu.py -q --tb=short > $O/pytest_encoder.log 2>&1; tail -3 $O/pytest_encoder.log | cut -c1-300
timeout 300 python tools/encoder_bench.py --aggregator mlp_mean > $O/encoder_mlp_mean.json 2> $O/encoder.err; cut -c1-1300 $O/encoder_mlp_mean.json; tail -2 $O/encoder.err
HOLO_VIEWPOOL_FUSE_ACT=0 timeout 300 python tools/encoder_bench.py --aggregator mlp_mean > $O/encoder_mlp_mean_unfused.json 2>> $O/encoder.err; cut -c1-900 $O/encoder_mlp_mean_unfused.json
timeout 300 python tools/encoder_bench.py --aggregator angle > $O/encoder_angle.json 2>> $O/encoder.err; cut -c1-900 $O/encoder_angle.json
timeout 300 python tools/encoder_bench.py --aggregator mlp_mean --views 4 > $O/encoder_mlp_mean_4views.json 2>> $O/encoder.err; cut -c1-900 $O/encoder_mlp_mean_4views.json
timeout 300 python tests/diagnostics/encoder_eager_compare.py > $O/encoder_eager_compare.json 2> $O/encoder_eager_compare.err; cat $O/encoder_eager_compare.json; tail -2 $O/encoder_eager_compare.err
