#!/bin/bash
# Round-2 measurement pass on one B200 (under gpurun): bench lines of every configuration, ncu launch lists, ncu --set full captures of the
# dominant kernels, compute-sanitizer. Outputs land in gpurun_out/ and are summarised into profiles/ by summarize_ncu.py r2.
set -u
O=gpurun_out
python bench.py > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err
python bench.py --impl reference > $O/r2_bench_ref.json 2> $O/r2_bench_ref.err
for c in 3 4 5; do python bench.py --config $c > $O/r2_bench_cfg$c.json 2> $O/r2_bench_cfg$c.err; done
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 1300 --csv --log-file $O/launches.csv python profiles/prof_run.py 1 2 > $O/r2_prof.log 2>&1
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/launches_cfg4.csv python profiles/prof_run.py 1 4 >> $O/r2_prof.log 2>&1
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/launches_cfg3.csv python profiles/prof_run.py 1 3 >> $O/r2_prof.log 2>&1
$NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $O/launches_cfg5.csv python profiles/prof_run.py 1 5 >> $O/r2_prof.log 2>&1
$NCU --set full --import-source on -k regex:conv_tower_kernel -s 30 -c 1 -f -o $O/tower_full python profiles/prof_run.py 1 2 >> $O/r2_prof.log 2>&1
$NCU --set full --import-source on -k regex:conv_tower_wide -s 30 -c 1 -f -o $O/tower_cfg4_full python profiles/prof_run.py 1 4 >> $O/r2_prof.log 2>&1
$NCU --set full --import-source on -k regex:k_step -s 60 -c 1 -f -o $O/kstep_full python profiles/prof_run.py 1 2 >> $O/r2_prof.log 2>&1
$NCU --set full --import-source on -k regex:heads_kernel -s 60 -c 1 -f -o $O/heads_full python profiles/prof_run.py 1 2 >> $O/r2_prof.log 2>&1
$NCU --set full --import-source on -k regex:conv_tower_kernel -s 40 -c 1 -f -o $O/tower_cfg3_full python profiles/prof_run.py 1 3 >> $O/r2_prof.log 2>&1
$NCU --set full --import-source on -k regex:conv_tower_kernel -s 40 -c 1 -f -o $O/tower_cfg5_full python profiles/prof_run.py 1 5 >> $O/r2_prof.log 2>&1
for tool in memcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python profiles/sanitizer_run.py > $O/r2_san_$tool.log 2>&1; tail -2 $O/r2_san_$tool.log
done
ls -la $O | tail -30
