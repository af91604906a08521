#!/bin/bash
# A/B: heads kernel as a programmatic dependent of the tower (per-board counter waits, CTAs fill the SMs the tower's last pass leaves idle) vs an ordinary
# kernel boundary (MZ_HEADS_OVERLAP=0; experiment build). usage (under gpurun): bash profiles/r2_heads_ab.sh > gpurun_out/r2_heads_ab.log 2>&1
set -u
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_shapes.py tests/test_gpu_think.py -x -q 2>&1 | tail -3
for rep in 1 2; do
for set in "MZ_HEADS_OVERLAP=0" "MZ_HEADS_OVERLAP=1"; do
  env KT_CONFIG=2 $set timeout 300 python profiles/kernel_times.py 2>&1 | tail -1
done
done
for set in "MZ_HEADS_OVERLAP=0" "MZ_HEADS_OVERLAP=1"; do
  env KT_CONFIG=4 $set timeout 300 python profiles/kernel_times.py 2>&1 | tail -1
done
