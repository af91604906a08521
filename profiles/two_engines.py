"""Experiment: 256 games as k interleaved sub-batches (k engines of 256/k games on one GPU, each with its own stream and
CUDA graph): while one sub-batch's tower keeps the tensor pipes busy, the other sub-batches' heads + tree steps run beside it —
on the SAME SMs when the tower leaves enough shared memory for them (MZ_TOWER_STAGES=4|5), else only on SMs the tower does
not occupy (MZ_TOWER_SMS)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import minizero_b200  # noqa: E402

n_eng = int(sys.argv[1]) if len(sys.argv) > 1 else 2
games = bench.GAMES // n_eng
engs = []
for i in range(n_eng):
    e = minizero_b200.Engine(minizero_b200.GAME_GO, bench.BOARD, games, bench.SIMS)
    e.load_network(bench.NET)
    engs.append(e)
if os.environ.get("PROFILE_KERNELS"):
    engs[0].set_search_inputs(None, None)
    engs[0].search()
    p = engs[0].profile_kernels(50)
    print(f"isolated kernels at {games} games: tower {p['conv_ms'] * 1e3:.1f} us  tree {p['tree_ms'] * 1e3:.1f} us  heads {p['heads_ms'] * 1e3:.1f} us")
    engs[0].reset_game(-1)
rng = np.random.default_rng(0)
rot = rng.integers(0, 8, size=(bench.SIMS + 1, games)).astype(np.uint8)
noise = rng.dirichlet([0.03] * bench.ACTIONS, size=games).astype(np.float32)
for rep in range(3):
    for e in engs:
        e.set_search_inputs(rot, noise)
    for e in engs:
        e.sync()
    t0 = time.perf_counter()
    for e in engs:
        e.search(wait=False)
    for e in engs:
        e.sync()
    dt = time.perf_counter() - t0
    for e in engs:
        e.play_max_count(auto_reset=True, read_back=True)
    tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("MZ_"))
    print(f"engines {n_eng} x {games} games: search {dt * 1e3:.1f} ms -> {bench.GAMES * (bench.SIMS + 1) / dt:.0f} evals/s  [{tag}]")
