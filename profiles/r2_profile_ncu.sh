#!/bin/bash
# ncu part of the round-2 measurement pass (launch lists + --set full captures of the towers), see r2_profile.sh
set -u
O=gpurun_out
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 1300 --csv --log-file $O/launches.csv python profiles/prof_run.py 1 2 > $O/r2_prof.log 2>&1
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/launches_cfg4.csv python profiles/prof_run.py 1 4 >> $O/r2_prof.log 2>&1
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/launches_cfg3.csv python profiles/prof_run.py 1 3 >> $O/r2_prof.log 2>&1
$NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $O/launches_cfg5.csv python profiles/prof_run.py 1 5 >> $O/r2_prof.log 2>&1
$NCU --set full --import-source on -k regex:conv_tower_kernel -s 30 -c 1 -f -o $O/tower_full python profiles/prof_run.py 1 2 >> $O/r2_prof.log 2>&1
$NCU --set full --import-source on -k regex:conv_tower_wide -s 30 -c 1 -f -o $O/tower_cfg4_full python profiles/prof_run.py 1 4 >> $O/r2_prof.log 2>&1
$NCU --set full --import-source on -k regex:conv_tower_kernel -s 40 -c 1 -f -o $O/tower_cfg3_full python profiles/prof_run.py 1 3 >> $O/r2_prof.log 2>&1
$NCU --set full --import-source on -k regex:conv_tower_kernel -s 40 -c 1 -f -o $O/tower_cfg5_full python profiles/prof_run.py 1 5 >> $O/r2_prof.log 2>&1
grep -c conv_tower $O/launches.csv $O/launches_cfg4.csv $O/launches_cfg3.csv $O/launches_cfg5.csv
timeout 600 python -m pytest tests/test_gpu_baseline_shapes.py tests/test_gpu_think.py tests/test_gpu_parity.py -x -q -k "cooperative or think or muzero" 2>&1 | tail -3
python bench.py --config 3 --no-cpu-baseline --no-gpu-reference 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg3', d['value'], d['e2e']['value'], d['roofline']['launch_ms'], d['kernels_ms'])"
