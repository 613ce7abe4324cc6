#!/bin/bash
# Round-end check on one box: the whole -m gpu suite, smoke, the default bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4 ) > gpurun_out/gputest.log 2>&1
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) > gpurun_out/smoke.log 2>&1
( timeout 600 python bench.py 2>&1 | tail -1 ) > gpurun_out/bench.log 2>&1
cat gpurun_out/gputest.log gpurun_out/smoke.log; tail -c 1500 gpurun_out/bench.log
