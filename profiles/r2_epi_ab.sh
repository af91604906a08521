#!/bin/bash
# A/B of the towers' new epilogue (residual rows fetched before the accumulators are ready, fp16 rows staged in shared memory and written by TMA stores,
# one red.release per warp instead of a __threadfence by every lane; wide tower: row tile 0 of a unit completes nine taps before tile 1).
# Needs a library built with MZ_BUILD_EXPERIMENT=1 (MZ_TOWER_WIDE forces the variant).
# usage (under gpurun): bash profiles/r2_epi_ab.sh > gpurun_out/r2_epi_ab.log 2>&1
set -u
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for set in "MZ_TOWER_WIDE=0" "MZ_TOWER_WIDE=0 MZ_DEBUG_TOWER=1" "MZ_TOWER_WIDE=1" "MZ_TOWER_WIDE=1 MZ_DEBUG_TOWER=1"; do
  env KT_CONFIG=2 $set timeout 300 python profiles/kernel_times.py 2>&1 | tail -3 | grep -v peers
done
for set in "MZ_TOWER_WIDE=1" "MZ_TOWER_WIDE=0"; do
  env KT_CONFIG=4 $set timeout 300 python profiles/kernel_times.py 2>&1 | tail -1
done
for cfg in 3 5; do
  for set in "MZ_TOWER_WIDE=0" "MZ_TOWER_WIDE=1"; do
    env $set timeout 300 python bench.py --config $cfg --no-cpu-baseline --no-gpu-reference 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg$cfg $set', d['value'], d['e2e']['value'], d['roofline']['launch_ms'], d['roofline']['frac'], d.get('kernels_ms'))"
  done
done
