#!/bin/bash
# release library: the whole GPU test suite, then the bench line of every configuration
# usage (under gpurun): bash profiles/r2_final_pass.sh > gpurun_out/r2_final_pass.log 2>&1
set -u
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
timeout 600 python bench.py > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; cat gpurun_out/bench_cfg2.json
for cfg in 3 4 5; do
  timeout 600 python bench.py --config $cfg --no-cpu-baseline --no-gpu-reference > gpurun_out/bench_cfg$cfg.json 2> gpurun_out/bench_cfg$cfg.err; cat gpurun_out/bench_cfg$cfg.json
done
