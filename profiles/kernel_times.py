"""Prints the CUDA-event time of the conv / tree / heads kernels on BASELINE configs[1] state (env MZ_CONV_* select the variant)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import minizero_b200  # noqa: E402

eng = minizero_b200.Engine(minizero_b200.GAME_GO, bench.BOARD, bench.GAMES, bench.SIMS)
eng.load_network(bench.NET)
rng = np.random.default_rng(0)
rot = rng.integers(0, 8, size=(bench.SIMS + 1, bench.GAMES)).astype(np.uint8)
noise = rng.dirichlet([0.03] * bench.ACTIONS, size=bench.GAMES).astype(np.float32)
eng.set_search_inputs(rot, noise)
ms = eng.search()
p = eng.profile_kernels(100)
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("MZ_"))
print(f"[{tag}] search {ms:.1f} ms ({bench.GAMES * (bench.SIMS + 1) / ms * 1e3:.0f} evals/s)  conv {p['conv_ms'] * 1e3:.1f} us  tree {p['tree_ms'] * 1e3:.1f} us  heads {p['heads_ms'] * 1e3:.1f} us")

