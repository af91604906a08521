#!/bin/bash
# A/B of the wide tower (two row tiles per CTA; chosen by shape, MZ_TOWER_WIDE=0|1 forces it in experiment builds) against the narrow one:
# usage (under gpurun): bash profiles/r2_wide_ab.sh > gpurun_out/r2_wide_ab.log 2>&1
set -u
echo "== release build (wide chosen by shape) =="
for cfg in 2 4; do
  KT_CONFIG=$cfg timeout 300 python profiles/kernel_times.py 2>&1 | tail -1
done
timeout 900 python -m pytest tests/test_gpu_baseline_shapes.py tests/test_gpu_parity.py -x -q 2>&1 | tail -5
timeout 300 python bench.py --config 3 --no-cpu-baseline --no-gpu-reference 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg3', d['value'], d['roofline']['launch_ms'], d['roofline']['frac'])"
timeout 300 python bench.py --config 5 --no-cpu-baseline --no-gpu-reference 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg5', d['value'], d['roofline']['launch_ms'], d['roofline']['frac'])"
cp minizero_b200/lib/libmzb200.so /tmp/libmzb200_release.so
MZ_BUILD_EXPERIMENT=1 python -c "import minizero_b200; minizero_b200.build_library(force=True)"
echo "== experiment build =="
for set in "MZ_TOWER_WIDE=0" "MZ_TOWER_WIDE=1" "MZ_TOWER_WIDE=1 MZ_TOWER_ROT=22" "MZ_TOWER_WIDE=1 MZ_TOWER_ROT=44" "MZ_TOWER_WIDE=1 MZ_TOWER_ROT=52" "MZ_TOWER_WIDE=1 MZ_TOWER_ROT=0"; do
  env KT_CONFIG=2 $set timeout 300 python profiles/kernel_times.py 2>&1 | tail -1
done
for set in "MZ_TOWER_WIDE=0" "MZ_TOWER_WIDE=0 MZ_TOWER_ROT=22" "MZ_TOWER_WIDE=1" "MZ_TOWER_WIDE=1 MZ_TOWER_ROT=30"; do
  env KT_CONFIG=4 $set timeout 300 python profiles/kernel_times.py 2>&1 | tail -1
done
env KT_CONFIG=2 MZ_DEBUG_TOWER=1 timeout 300 python profiles/kernel_times.py 2>&1 | tail -3
env KT_CONFIG=4 MZ_DEBUG_TOWER=1 timeout 300 python profiles/kernel_times.py 2>&1 | tail -3
cp /tmp/libmzb200_release.so minizero_b200/lib/libmzb200.so
