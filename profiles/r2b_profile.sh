#!/bin/bash
# Measurement pass after the towers' epilogues went to TMA stores (tag r2b; one B200, under gpurun): ncu launch lists, ncu --set full captures of the
# towers, compute-sanitizer. Outputs land in gpurun_out/ and are summarised into profiles/ by `python profiles/summarize_ncu.py r2b`.
set -u
O=gpurun_out
rm -f $O/launches*.csv $O/tower*_full.ncu-rep
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 1300 --csv --log-file $O/launches.csv python profiles/prof_run.py 1 2 > $O/r2b_prof.log 2>&1
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/launches_cfg4.csv python profiles/prof_run.py 1 4 >> $O/r2b_prof.log 2>&1
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/launches_cfg3.csv python profiles/prof_run.py 1 3 >> $O/r2b_prof.log 2>&1
$NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $O/launches_cfg5.csv python profiles/prof_run.py 1 5 >> $O/r2b_prof.log 2>&1
$NCU --set full --import-source on -k regex:conv_tower -s 30 -c 1 -f -o $O/tower_full python profiles/prof_run.py 1 2 >> $O/r2b_prof.log 2>&1
$NCU --set full --import-source on -k regex:conv_tower -s 30 -c 1 -f -o $O/tower_cfg4_full python profiles/prof_run.py 1 4 >> $O/r2b_prof.log 2>&1
$NCU --set full --import-source on -k regex:conv_tower -s 40 -c 1 -f -o $O/tower_cfg3_full python profiles/prof_run.py 1 3 >> $O/r2b_prof.log 2>&1
$NCU --set full --import-source on -k regex:conv_tower -s 40 -c 1 -f -o $O/tower_cfg5_full python profiles/prof_run.py 1 5 >> $O/r2b_prof.log 2>&1
grep -c conv_tower $O/launches.csv $O/launches_cfg4.csv $O/launches_cfg3.csv $O/launches_cfg5.csv
for tool in memcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool python profiles/sanitizer_run.py > $O/r2b_san_$tool.log 2>&1; tail -3 $O/r2b_san_$tool.log
done
ls -la $O | tail -12
