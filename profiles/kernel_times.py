"""Prints the CUDA-event time of the conv / tree / heads kernels on the state of a BASELINE configuration (KT_CONFIG = bench.py's --config
numbering, default 2). The MZ_* experiment switches need a library built with MZ_BUILD_EXPERIMENT=1."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import minizero_b200  # noqa: E402

w = bench.WORKLOADS[int(os.environ.get("KT_CONFIG", "2"))]
GAMES, SIMS = w["games"], w["sims"]
eng = minizero_b200.Engine(minizero_b200.GAME_GO, w["board"], GAMES, SIMS)
path = os.path.join(bench.NETS, w["net"] + ".pt")
if os.path.exists(path):
    eng.load_network(path)
else:
    import __graft_entry__ as ge
    eng.load_network((w["dims"], ge.make_random_state(w["dims"], np.random.default_rng(0))))
rng = np.random.default_rng(0)
rot = rng.integers(0, 8, size=(SIMS + 1, GAMES)).astype(np.uint8)
noise = rng.dirichlet([0.03] * eng.A, size=GAMES).astype(np.float32)
eng.set_search_inputs(rot, noise)
ms = eng.search()
p = eng.profile_kernels(100)
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("MZ_") or k.startswith("KT_"))
print(f"[{tag}] search {ms:.1f} ms ({GAMES * (SIMS + 1) / ms * 1e3:.0f} evals/s)  conv {p['conv_ms'] * 1e3:.1f} us  tree {p['tree_ms'] * 1e3:.1f} us  heads {p['heads_ms'] * 1e3:.1f} us")


if os.environ.get("MZ_DEBUG_TREE"):
    eng.tree_timing()  # reset the counters
    eng.play_max_count(auto_reset=True, read_back=True)
    eng.set_search_inputs(rot, noise)
    ms = eng.search()
    t = eng.tree_timing().astype(np.float64)
    steps = t[:, 5].max()
    names = ["select", "transition", "analysis", "features", "expand+backup"]
    per_step = t[:, :5] / np.maximum(t[:, 5:6], 1)
    print("in-situ search %.1f ms; tree cycles per step per game (mean over games | max over games):" % ms)
    for i, n in enumerate(names):
        print("  %-14s %8.0f | %8.0f" % (n, per_step[:, i].mean(), per_step[:, i].max()))
    print("  path length mean %.1f, longest %d" % ((t[:, 7] / np.maximum(t[:, 5], 1)).mean(), t[:, 6].max()))
    st = np.maximum(t[:, 5:6], 1)
    d = t[:, 8:14] / st
    print("  select detail per step (mean over games): check %.0f cyc, chase %.0f cyc, serial finish %.0f cyc; rounds %.2f, levels checked %.1f, serial levels %.2f" % tuple(d.mean(axis=0)))

if os.environ.get("MZ_DEBUG_TOWER"):
    eng.tower_timing()
    t = eng.tower_timing().astype(np.float64)
    names = ["prod_total", "prod_wait_deps", "prod_wait_stage", "mma_total", "mma_wait_block", "mma_wait_acc", "mma_wait_weights", "epi_busy"]
    lead = t[0::2]
    print("tower per-CTA cycles, leaders (mean / max):", ", ".join(f"{n}={lead[:, i].mean():.0f}/{lead[:, i].max():.0f}" for i, n in enumerate(names)))
    peer = t[1::2]
    print("tower per-CTA cycles, peers   (mean / max):", ", ".join(f"{n}={peer[:, i].mean():.0f}/{peer[:, i].max():.0f}" for i, n in enumerate(names)))
