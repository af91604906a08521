#!/bin/bash
# A/B of the wide tower's shared-memory split: input ring of 2 K-blocks + 14 weight stages (this build) against ring 3 + 10 / 9 stages (profiles/r2_epi_ab.log)
set -u
timeout 600 python -m pytest tests/test_gpu_baseline_shapes.py -x -q -k "wide_and_narrow or 20bx256 or saturates or trained_scale or config3" 2>&1 | tail -2
for rep in 1 2; do
for set in "KT_CONFIG=2" "KT_CONFIG=2 MZ_DEBUG_TOWER=1" "KT_CONFIG=4" "KT_CONFIG=4 MZ_DEBUG_TOWER=1"; do
  env $set timeout 300 python profiles/kernel_times.py 2>&1 | tail -3 | grep -v peers
done
done
