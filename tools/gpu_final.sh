#!/bin/bash
# Round-2 validation + evidence call (1 GPU): smoke, full GPU suite, the bench line, steady-state launch list with DRAM
# bytes, `ncu --set full` of the top kernels, the 1000-step chain drift with its fp32 yardstick.
O=gpurun_out/final; mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6) > $O/pytest_all.log 2>&1; tail -3 $O/pytest_all.log
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err; cut -c1-400 $O/bench.json; tail -2 $O/bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; cut -c1-300 $O/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv --log-file $O/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-eager-baseline --no-clocks > $O/ncu_bench.log 2>&1
NB="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-eager-baseline --no-clocks"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 260 -c 3 -o $O/prof_conv_tc -f $NB > $O/ncu_full.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:render_tc_kernel -s 3 -c 1 -o $O/prof_render_tc -f $NB >> $O/ncu_full.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_flash_kernel -s 31 -c 2 -o $O/prof_attn_flash -f $NB >> $O/ncu_full.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:splitk_reduce_kernel -s 120 -c 2 -o $O/prof_splitk_reduce -f $NB >> $O/ncu_full.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gn_apply_fused -s 200 -c 2 -o $O/prof_gn_apply -f $NB >> $O/ncu_full.log 2>&1
tail -3 $O/ncu_full.log
timeout 500 python tests/diagnostics/chain_drift.py --resol 32 --steps 1000 --every 100 --f64 --with-eager32 > $O/drift_32_1000_f64_eager32.json 2> $O/drift.err; cut -c1-300 $O/drift_32_1000_f64_eager32.json; tail -2 $O/drift.err
ls -la $O
