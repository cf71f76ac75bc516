#!/bin/bash
# Round-2 call 9 (8 GPUs of one box): scaling of the headline bench, cfg #4 (batch-32, one gather), cfg #5 (128^3, one
# sample on 8 GPUs: query-sharded attention + row-sharded render), multi-GPU parity at 32^3.
O=gpurun_out/c9; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name,memory.total --format=csv > $O/smi.txt 2>&1
timeout 300 $TR --master-port 29501 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_8gpu.json 2> $O/bench_8gpu.err
grep '^{' $O/bench_8gpu.json | cut -c1-330; tail -2 $O/bench_8gpu.err
timeout 300 $TR --master-port 29502 tools/batch_sharded.py --batch 32 --repeats 3 > $O/cfg4_8gpu.json 2> $O/cfg4_8gpu.err
grep '^{' $O/cfg4_8gpu.json; tail -2 $O/cfg4_8gpu.err
timeout 300 $TR --master-port 29503 tests/diagnostics/check_one_sample_multi_gpu.py --resol 32 --channels 16 --image 128 --pts 32 --attn-min-tokens 512 > $O/cfg5_parity_32_8gpu.json 2> $O/cfg5_parity.err
grep '^{' $O/cfg5_parity_32_8gpu.json; tail -2 $O/cfg5_parity.err
timeout 500 $TR --master-port 29504 tools/one_sample_multi_gpu.py --resol 128 --image 512 --pts 128 --fine 16 --steps 1 > $O/cfg5_128_8gpu.json 2> $O/cfg5_128.err
grep '^{' $O/cfg5_128_8gpu.json; tail -3 $O/cfg5_128.err
nvidia-smi --query-gpu=index,memory.used --format=csv >> $O/smi.txt 2>&1
