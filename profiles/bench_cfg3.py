"""BASELINE.json configs[2] (a parity-test configuration, not the bench line): 8x8 Othello Gumbel MuZero, n=16 simulations,
512 parallel games, 3-block x 128-channel network. Prints one JSON line: leaf-evals/s of the on-device search (CUDA events on
the engine's stream, inputs resident), the per-kernel times, and the unmodified reference ActorGroup on the host cores.

    python profiles/bench_cfg3.py [steps] [--no-ref]
"""
import json
import os
import re
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import minizero_b200  # noqa: E402

GAMES, SIMS, N, M = 512, 16, 8, 16
NET = os.path.join(ROOT, "oracle", "_ref", "nets", "othello_mz_3bx128.pt")
# 0.1139 GFLOP initial inference, 0.1324 GFLOP recurrent inference per position
FLOPS_PER_MOVE_PER_GAME = 0.1139e9 + SIMS * 0.1324e9


def reference(cycles):
    binary = os.path.join(ROOT, "oracle", "_ref", "ref_actor_group_othello")
    if not os.path.exists(binary):
        return None
    cores = os.cpu_count() or 1
    conf = (f"actor_num_simulation={SIMS}:zero_num_parallel_games={GAMES}:zero_num_threads={cores}:nn_type_name=muzero:actor_use_gumbel=true:"
            f"actor_use_gumbel_noise=true:actor_gumbel_sample_size={M}:actor_gumbel_sigma_visit_c=50:actor_gumbel_sigma_scale_c=1:"
            f"actor_use_dirichlet_noise=false:nn_file_name={NET}:program_seed=1:program_auto_seed=false:program_quiet=true")
    env = dict(os.environ, OMP_NUM_THREADS=str(cores), CUDA_VISIBLE_DEVICES="")
    res = subprocess.run([binary, "bench", conf, "4", str(cycles), "-1"], capture_output=True, text=True, env=env, timeout=900)
    m = re.search(r"REFBENCH evals=(\d+) seconds=([0-9.eE+-]+)", res.stdout)
    if not m:
        return {"error": (res.stdout + res.stderr)[-300:]}
    return {"value": int(m.group(1)) / float(m.group(2)), "unit": "leaf-evals/s", "cores": cores, "kind": "reference",
            "sample": f"{cycles} ActorGroup cycles x {GAMES} games, all-CPU"}


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 50
    eng = minizero_b200.Engine(minizero_b200.GAME_OTHELLO, N, GAMES, SIMS, muzero=1, use_gumbel=1, gumbel_noise=1, gumbel_sample_size=M)
    eng.load_network(NET)
    rng = np.random.default_rng(0)
    noise = rng.gumbel(size=(GAMES, eng.A)).astype(np.float32)
    eng.set_search_inputs(None, noise)
    for _ in range(5):
        eng.search(wait=False)
        eng.play_max_count(auto_reset=True, read_back=False)
    eng.sync()
    eng.timer_begin()
    for _ in range(steps):
        eng.search(wait=False)
        eng.play_max_count(auto_reset=True, read_back=False)
    ms = eng.timer_end()
    evals = GAMES * (SIMS + 1) * steps
    prof = eng.profile_kernels(50)
    line = {"workload": "othello8x8_gumbel_muzero_n16_512games_3bx128 (BASELINE configs[2])", "metric": "selfplay_leaf_evals_per_sec",
            "value": evals / (ms * 1e-3), "unit": "leaf-evals/s", "steps": steps, "ms_per_step": ms / steps,
            "moves_per_sec": GAMES * steps / (ms * 1e-3), "tflops_algorithmic": GAMES * steps * FLOPS_PER_MOVE_PER_GAME / (ms * 1e-3) / 1e12,
            "kernels_ms": {"dynamics_tower": prof["conv_ms"], "tree_step": prof["tree_ms"], "heads": prof["heads_ms"]},
            "gpu_launches_per_step": 1 + (SIMS + 1) * 4 + 1}
    if "--no-ref" not in sys.argv:
        line["cpu_baseline"] = reference(8)
    print(json.dumps(line))


if __name__ == "__main__":
    main()
