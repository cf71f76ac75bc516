#!/bin/bash
# End-of-round validation (1 GPU), what the driver runs: smoke, the full GPU suite, the bench line.
O=gpurun_out/final2; mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6) > $O/pytest_all.log 2>&1; tail -3 $O/pytest_all.log
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err; cut -c1-500 $O/bench.json; tail -2 $O/bench.err
