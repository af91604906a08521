#!/bin/bash
# A/B of the tower experiment switches on one box: experiment build (knobs compiled in), CUDA-event kernel times per setting.
# usage (under gpurun): bash profiles/r2_tower_ab.sh > gpurun_out/r2_tower_ab.log 2>&1
set -u
cp minizero_b200/lib/libmzb200.so /tmp/libmzb200_release.so
MZ_BUILD_EXPERIMENT=1 python -c "import minizero_b200; minizero_b200.build_library(force=True)"
for cfg in 2 4; do
  for set in "MZ_TOWER_BN=128" "MZ_TOWER_BN=256" "MZ_TOWER_BN=256 MZ_TOWER_FENCE=1"; do
    env KT_CONFIG=$cfg $set python profiles/kernel_times.py 2>&1 | tail -1
  done
done
env MZ_TOWER_BN=256 python -m pytest tests/test_gpu_parity.py -q -x -k "go9_az_6bx256 or full_size_search or go9_deep" 2>&1 | tail -3
cp /tmp/libmzb200_release.so minizero_b200/lib/libmzb200.so
