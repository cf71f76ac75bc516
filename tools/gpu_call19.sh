#!/bin/bash
# encoder: embedding column per lane; ncu capture of the gather / elementwise kernels
O=gpurun_out/c19; mkdir -p $O
timeout 600 python -m pytest tests/test_encoder_gpu.py -q -s --tb=short > $O/pytest_encoder.log 2>&1; grep -v "^$" $O/pytest_encoder.log | tail -14 | cut -c1-300
timeout 300 python tools/encoder_bench.py --aggregator mlp_mean --chunk 32768 > $O/encoder_mlp_mean_32768.json 2> $O/encoder.err; cut -c1-1300 $O/encoder_mlp_mean_32768.json; tail -2 $O/encoder.err
timeout 300 python tools/encoder_bench.py --aggregator angle --chunk 32768 > $O/encoder_angle_32768.json 2>> $O/encoder.err; cut -c1-900 $O/encoder_angle_32768.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:viewpool -c 4 -f -o $O/viewpool python tools/encoder_bench.py --aggregator mlp_mean --chunk 32768 --iters 1 > $O/ncu.log 2>&1; tail -3 $O/ncu.log
ls -la $O/*.ncu-rep
