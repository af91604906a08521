#!/bin/bash
# 8-GPU pass of round 2 (under `gpurun --gpus 8`): the drop-in executable on 4 / 8 GPUs (one host thread per engine), the bench on 8 GPUs for configs 2, 3, 4
set -u
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python profiles/worker_throughput.py 25 4 > $O/r2_worker_throughput_n4.json 2> $O/wt4.err
python profiles/worker_throughput.py 25 8 > $O/r2_worker_throughput_n8.json 2> $O/wt8.err
$TR --nproc-per-node 8 --master-port 29511 bench.py --gpus 8 --no-cpu-baseline --no-gpu-reference > $O/r2_bench_n8.json 2> $O/r2_bench_n8.err
$TR --nproc-per-node 8 --master-port 29512 bench.py --gpus 8 --config 3 --no-cpu-baseline --no-gpu-reference > $O/r2_bench_cfg3_n8.json 2> $O/r2_bench_cfg3_n8.err
$TR --nproc-per-node 8 --master-port 29513 bench.py --gpus 8 --config 4 --steps 4 --no-cpu-baseline --no-gpu-reference > $O/r2_bench_cfg4_n8.json 2> $O/r2_bench_cfg4_n8.err
timeout 300 python -m pytest tests/test_gpu_worker.py -q -k "all_visible_gpus" 2>&1 | tail -2
for f in r2_worker_throughput_n4 r2_worker_throughput_n8 r2_bench_n8 r2_bench_cfg3_n8 r2_bench_cfg4_n8; do echo $f; python -c "
import json
d=json.loads(open('$O/$f.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('n_gpus','value','ms_per_step','leaf_evals_per_sec_steady','ms_per_engine_move','engine_moves')}, (d.get('e2e') or {}).get('value'))"; done
