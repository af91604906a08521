"""Experiment: 256 games as two interleaved half-batches (two engines of 128 games on one GPU, each with its own stream and
CUDA graph): while one half's tower occupies most SMs, the other half's heads + tree step run on the SMs left free."""
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
    print(f"engines {n_eng} x {games} games: search {dt * 1e3:.1f} ms -> {bench.GAMES * (bench.SIMS + 1) / dt:.0f} evals/s  [MZ_TOWER_SMS={os.environ.get('MZ_TOWER_SMS')}]")
