#!/bin/bash
# Run on the GPU box (under gpurun): launch list (+ optionally full captures of the dominant kernels) of one bench step.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
if [ "$1" == "full" ]; then
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 1 -c 2 -o gpurun_out/prof_conv_tc -f \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:render_tc_kernel -s 1 -c 1 -o gpurun_out/prof_render_tc -f \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline >> gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_flash_kernel -s 11 -c 2 -o gpurun_out/prof_attn_flash -f \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline >> gpurun_out/ncu_full.log 2>&1
fi
ls -la gpurun_out
