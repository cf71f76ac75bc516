#!/bin/bash
# Round-2 call 5: split-K statistics by the last slice, new bench instrumentation + eager comparator + ViewStream e2e,
# conv launch anatomy, chain drift with the fp32 eager yardstick.
O=gpurun_out/c5; mkdir -p $O
(timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15) > $O/pytest_all.log 2>&1
tail -4 $O/pytest_all.log
timeout 400 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err; cut -c1-300 $O/bench.json; tail -2 $O/bench.err
timeout 300 python tools/conv_trace.py > $O/conv_trace.log 2>&1; cat $O/conv_trace.log
HOLO_PDL=1 timeout 300 python tools/conv_trace.py --no-trace > $O/conv_trace_pdl.log 2>&1; cat $O/conv_trace_pdl.log
timeout 400 python tests/diagnostics/chain_drift.py --resol 32 --steps 250 --every 25 --f64 --with-eager32 > $O/drift_32_250_f64.json 2> $O/drift.err; cut -c1-300 $O/drift_32_250_f64.json; tail -2 $O/drift.err
