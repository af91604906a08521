#!/bin/bash
# wide tower: completion counters per block of 64 output channels (= per K-block of the next layer), published as soon as both row tiles of the block are stored
set -u
timeout 900 python -m pytest tests/test_gpu_baseline_shapes.py tests/test_gpu_think.py -x -q 2>&1 | tail -2
for rep in 1 2; do
for set in "KT_CONFIG=2" "KT_CONFIG=2 MZ_DEBUG_TOWER=1" "KT_CONFIG=4"; do
  env $set timeout 300 python profiles/kernel_times.py 2>&1 | tail -3 | grep -v peers
done
done
timeout 300 python bench.py --config 3 --no-cpu-baseline --no-gpu-reference 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg3', d['value'], d['e2e']['value'], d['roofline']['launch_ms'], d['roofline']['frac'])"
