#!/bin/bash
# warps per game in the tree-step kernel: 16 (64 registers per thread at two blocks per SM: 380 bytes of spill loads) against 12 (up to 80 registers)
# and 8 (128); CUDA-event kernel times at configs 2 and 4
set -u
cp minizero_b200/lib/libmzb200.so /tmp/libmzb200_release.so
cp minizero_b200/csrc/engine.cu /tmp/engine.cu.orig
for w in 16 12 8; do
  sed "s/constexpr int STEP_WARPS = 16;/constexpr int STEP_WARPS = $w;/" /tmp/engine.cu.orig > minizero_b200/csrc/engine.cu
  python -c "import minizero_b200; minizero_b200.build_library(force=True, verbose=True)" 2>&1 | grep -A2 "Function properties for _ZN41.*k_stepE7" | grep -E "spill|registers" | tr '\n' ' '
  echo "  <- STEP_WARPS=$w"
  for cfg in 2 4; do KT_CONFIG=$cfg timeout 300 python profiles/kernel_times.py 2>&1 | tail -1; done
done
cp /tmp/engine.cu.orig minizero_b200/csrc/engine.cu
cp /tmp/libmzb200_release.so minizero_b200/lib/libmzb200.so
