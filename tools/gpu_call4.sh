#!/bin/bash
# Round-2 call 4: full GPU suite (post-processing kernels, shim fly-around), bench with balanced chunk 9, steady-state
# ncu launch list with DRAM bytes, DDPM chain drift.
O=gpurun_out/c4; mkdir -p $O
(timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15) > $O/pytest_all.log 2>&1
tail -4 $O/pytest_all.log
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
$B > $O/bench.json 2> $O/bench.err; cut -c1-300 $O/bench.json; tail -2 $O/bench.err
timeout 400 python tests/diagnostics/chain_drift.py --resol 32 --steps 1000 --f64 > $O/drift_32_1000_f64.json 2> $O/drift.err; cut -c1-400 $O/drift_32_1000_f64.json
timeout 300 python tests/diagnostics/chain_drift.py --resol 64 --steps 100 --every 10 > $O/drift_64_100_f32.json 2>> $O/drift.err; cut -c1-400 $O/drift_64_100_f32.json
tail -3 $O/drift.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv --log-file $O/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-clocks > $O/ncu_bench.log 2>&1
ls -la $O | head -20
