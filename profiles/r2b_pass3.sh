#!/bin/bash
# after the epilogue choice per tower (TMA stores where a layer's hand-off is on the critical path, lane-per-row stores for the many-unit Atari stages):
# GPU tests, config-5 bench line, fresh launch lists of configs 2 / 3 / 5
set -u
O=gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > $O/r2b_pytest_gpu.log 2>&1; tail -2 $O/r2b_pytest_gpu.log
timeout 900 python bench.py --config 5 > $O/r2b_bench_cfg5.json 2> $O/r2b_bench_cfg5.err; cut -c1-200 $O/r2b_bench_cfg5.json
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 1300 --csv --log-file $O/launches.csv python profiles/prof_run.py 1 2 > $O/r2b_prof.log 2>&1
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/launches_cfg3.csv python profiles/prof_run.py 1 3 >> $O/r2b_prof.log 2>&1
$NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $O/launches_cfg5.csv python profiles/prof_run.py 1 5 >> $O/r2b_prof.log 2>&1
grep -c conv_tower $O/launches.csv $O/launches_cfg3.csv $O/launches_cfg5.csv
