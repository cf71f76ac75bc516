#!/bin/bash
O=gpurun_out/c11; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tests/diagnostics/check_one_sample_multi_gpu.py --resol 32 --channels 16 --image 128 --pts 32 --attn-min-tokens 512 > $O/cfg5_parity_32_2gpu.json 2> $O/cfg5_parity.err
grep '^{' $O/cfg5_parity_32_2gpu.json; tail -2 $O/cfg5_parity.err
timeout 300 $TR --master-port 29512 tools/one_sample_multi_gpu.py --resol 64 --image 256 --pts 64 --steps 1 --profile-attention > $O/cfg5_64_2gpu.json 2> $O/cfg5_64.err
grep '^{' $O/cfg5_64_2gpu.json; tail -2 $O/cfg5_64.err
(timeout 300 python -m pytest tests/test_sharded_attention_gpu.py tests/test_unet_gpu.py -q -x -k "sharded or every_level" 2>&1 | tail -3) > $O/pytest.log 2>&1; cat $O/pytest.log
