#!/bin/bash
# Round-end check on one box: the whole -m gpu suite, smoke, the default bench line, launch lists of one forward per config
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -8 ) > gpurun_out/gputest.log 2>&1
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) > gpurun_out/smoke.log 2>&1
( timeout 600 python bench.py 2>&1 | tail -3 ) > gpurun_out/bench.log 2>&1
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 101 -c 101 --csv --log-file gpurun_out/launches_r2c.csv python tools/profile_forward.py --batch 256 --forwards 2 2>&1 | tail -2 ) > gpurun_out/launches.log 2>&1
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 101 -c 101 --csv --log-file gpurun_out/launches_deits_r2c.csv python tools/profile_forward.py --model deit_small_patch16_224 --batch 128 --forwards 2 2>&1 | tail -2 ) >> gpurun_out/launches.log 2>&1
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_swin_r2c.csv python tools/profile_swin.py 128 2>&1 | tail -2 ) >> gpurun_out/launches.log 2>&1
tail -4 gpurun_out/gputest.log gpurun_out/smoke.log gpurun_out/launches.log; tail -c 2500 gpurun_out/bench.log
