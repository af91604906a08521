#!/bin/bash
# release library at HEAD: GPU test suite, bench lines of every configuration (with both reference arms), racecheck, the config-3 tower capture,
# the drop-in executable's throughput on one GPU. usage (under gpurun): bash profiles/r2b_final.sh > gpurun_out/r2b_final.log 2>&1
set -u
O=gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > $O/r2b_pytest_gpu.log 2>&1; tail -2 $O/r2b_pytest_gpu.log
timeout 600 python bench.py > $O/r2b_bench_n1.json 2> $O/r2b_bench_n1.err; cut -c1-400 $O/r2b_bench_n1.json
timeout 600 python bench.py --impl reference > $O/r2b_bench_ref.json 2> $O/r2b_bench_ref.err; cut -c1-300 $O/r2b_bench_ref.json
for c in 3 4 5; do timeout 900 python bench.py --config $c > $O/r2b_bench_cfg$c.json 2> $O/r2b_bench_cfg$c.err; cut -c1-200 $O/r2b_bench_cfg$c.json; done
ncu --clock-control none --set full --import-source on -k regex:conv_tower -s 30 -c 1 -f -o $O/tower_full python profiles/prof_run.py 1 2 > $O/r2b_prof2.log 2>&1
ncu --clock-control none --set full --import-source on -k regex:conv_tower -s 8 -c 1 -f -o $O/tower_cfg3_full python profiles/prof_run.py 1 3 > $O/r2b_prof3.log 2>&1
for tool in memcheck synccheck; do timeout 900 compute-sanitizer --tool $tool python profiles/sanitizer_run.py > $O/r2b_san_$tool.log 2>&1; tail -1 $O/r2b_san_$tool.log; done
timeout 600 python profiles/worker_throughput.py > $O/r2b_worker_throughput_n1.json 2> $O/r2b_wt1.err; tail -c 600 $O/r2b_worker_throughput_n1.json
