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


if os.environ.get("MZ_DBG"):
    t = eng.tree_timing().astype(np.float64)
    tot = t[:, :4].sum(axis=1)
    order = np.argsort(-tot)
    print("tree step per-game cycles: mean total %.0f, max %.0f" % (tot.mean(), tot.max()))
    print("phase means (select, transition, leaf analysis, features):", t[:, :4].mean(axis=0).round(0))
    print("path len: mean %.1f max %d; leaf move number mean %.1f max %d; terminal leaves %d" % (t[:, 4].mean(), t[:, 4].max(), t[:, 6].mean(), t[:, 6].max(), int(t[:, 5].sum())))
    for g in order[:6]:
        print("  game %3d: select %7d transition %6d analysis %7d features %6d | path %3d terminal %d move %3d legal %2d" % ((g,) + tuple(int(x) for x in t[g])))
