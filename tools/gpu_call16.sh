#!/bin/bash
# validation after removing the unused TMEM helpers: ViewStream test with a traceback, whole GPU suite, smoke, bench
O=gpurun_out/c16; mkdir -p $O
timeout 300 python -m pytest tests/test_model_gpu.py -q -x -s --tb=short -k "view_stream" > $O/pytest_vs.log 2>&1; tail -30 $O/pytest_vs.log | cut -c1-300
timeout 900 python -m pytest tests -q -m gpu --tb=short > $O/pytest_all.log 2>&1; tail -8 $O/pytest_all.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err; cut -c1-600 $O/bench.json; tail -2 $O/bench.err
