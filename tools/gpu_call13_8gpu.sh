#!/bin/bash
# Round-2 call 13 (8 GPUs): cfg #5 at full size with the chunked-O attention, per-block attention / all-gather times;
# multi-GPU parity at 32^3; the headline bench at N = 8 once more (e2e through ViewStream).
O=gpurun_out/c13; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29521 tools/one_sample_multi_gpu.py --resol 128 --image 512 --pts 128 --fine 16 --steps 1 --profile-attention > $O/cfg5_128_8gpu.json 2> $O/cfg5_128.err
grep '^{' $O/cfg5_128_8gpu.json; tail -2 $O/cfg5_128.err
timeout 300 $TR --master-port 29522 tests/diagnostics/check_one_sample_multi_gpu.py --resol 32 --channels 16 --image 128 --pts 32 --attn-min-tokens 512 > $O/cfg5_parity_32_8gpu.json 2> $O/cfg5_parity.err
grep '^{' $O/cfg5_parity_32_8gpu.json; tail -2 $O/cfg5_parity.err
timeout 300 $TR --master-port 29523 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_8gpu.json 2> $O/bench_8gpu.err
grep '^{' $O/bench_8gpu.json | cut -c1-330; tail -2 $O/bench_8gpu.err
