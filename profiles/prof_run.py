"""Short driver for ncu: one whole-move search of BASELINE configs[1] (256 games, 400 sims, 6bx256) on cuda:0.
Usage under gpurun (see profiles/README.md): ncu ... python profiles/prof_run.py [num_searches]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import minizero_b200  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
eng = minizero_b200.Engine(minizero_b200.GAME_GO, bench.BOARD, bench.GAMES, bench.SIMS)
eng.load_network(bench.NET)
rng = np.random.default_rng(0)
rot = rng.integers(0, 8, size=(bench.SIMS + 1, bench.GAMES)).astype(np.uint8)
noise = rng.dirichlet([0.03] * bench.ACTIONS, size=bench.GAMES).astype(np.float32)
for i in range(n):
    eng.set_search_inputs(rot, noise)
    ms = eng.search()
    print("search", i, "ms", ms, "evals/s", bench.GAMES * (bench.SIMS + 1) / ms * 1e3)
    eng.play_max_count(auto_reset=True, read_back=True)
