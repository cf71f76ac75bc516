#!/bin/bash
# encoder: activations folded into the GEMM epilogues (holo_gemm_tc_act), work buffers kept between calls
O=gpurun_out/c20; mkdir -p $O
timeout 600 python -m pytest tests/test_encoder_gpu.py -q -s --tb=short > $O/pytest_encoder.log 2>&1; grep -v "^$" $O/pytest_encoder.log | tail -24 | cut -c1-300
for f in 1 0; do
  HOLO_VIEWPOOL_FUSE_ACT=$f timeout 300 python tools/encoder_bench.py --aggregator mlp_mean --chunk 32768 > $O/encoder_mlp_mean_fuse$f.json 2> $O/encoder_fuse$f.err; cut -c1-1300 $O/encoder_mlp_mean_fuse$f.json; tail -2 $O/encoder_fuse$f.err
done
timeout 300 python tools/encoder_bench.py --aggregator mlp_mean > $O/encoder_mlp_mean_default_chunk.json 2>> $O/encoder_fuse1.err; cut -c1-700 $O/encoder_mlp_mean_default_chunk.json
