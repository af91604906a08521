"""Short driver for ncu: whole-move searches of one BASELINE configuration (bench.py's WORKLOADS numbering; default 2 = Go 9x9,
256 games, 400 simulations, 6b x 256) on cuda:0.
Usage under gpurun (see profiles/README.md): ncu ... python profiles/prof_run.py [num_searches] [config]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import minizero_b200  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
w = bench.WORKLOADS[int(sys.argv[2]) if len(sys.argv) > 2 else 2]
games, sims = w["games"], w["sims"]
game = {"go": minizero_b200.GAME_GO, "othello": minizero_b200.GAME_OTHELLO, "atari": minizero_b200.GAME_ATARI}[w["game"]]
eng = minizero_b200.Engine(game, w["board"], games, sims, muzero=w["muzero"], **w.get("engine", {}))
path = os.path.join(bench.NETS, w["net"] + ".pt")
if os.path.exists(path):
    eng.load_network(path)
else:
    import __graft_entry__ as ge
    eng.load_network((w["dims"], ge.make_random_state(w["dims"], np.random.default_rng(0))))
rng = np.random.default_rng(0)
rot = None if w["muzero"] else rng.integers(0, 8, size=(sims + 1, games)).astype(np.uint8)
gumbel = bool(w.get("engine", {}).get("use_gumbel"))
noise = (rng.gumbel(size=(games, eng.A)) if gumbel else rng.dirichlet([0.03] * eng.A, size=games)).astype(np.float32)
if w["game"] == "atari":
    eng.observe_all(np.full(games, -1, np.int32), rng.integers(0, 256, size=(games, 3, 96, 96), dtype=np.uint8))
for i in range(n):
    eng.set_search_inputs(rot, noise)
    ms = eng.search()
    print("search", i, "ms", ms, "evals/s", games * (sims + 1) / ms * 1e3)
    eng.play_max_count(auto_reset=True, read_back=True)
