"""Differential fuzzing of two search engines (any pair of: CPU oracle, host-sim build of search_core.cuh, CUDA engine) that expose the
per-phase protocol select / apply / root / play: both are driven move after move with the SAME synthetic network outputs — random
priors (sometimes quantised so that exact ties occur), random values, root noise, rotations — and must agree on every leaf (path
length, feature planes) and on the root child table of every move, bit for bit. Independent of any recording: it reaches tree shapes
self-play with a fixed net does not (wide roots, nodes with many visited children, repeated terminal leaves, superko along paths)."""
import numpy as np


def run(a, b, *, num_actions, sims, games, moves, seed, rotations=True, noise="dirichlet", muzero=False, gumbel=False, check_features=True):
    rng = np.random.default_rng(seed)
    A, B, S = num_actions, games, sims
    compared = 0
    for move in range(moves):
        if noise == "dirichlet":
            nz = rng.dirichlet([0.3] * A, size=B).astype(np.float32)
        elif noise == "gumbel":
            nz = rng.gumbel(size=(B, A)).astype(np.float32)
        else:
            nz = None
        for c in range(S + 1):
            rot = rng.integers(0, 8, size=B).astype(np.uint8) if (rotations and not muzero) else None
            fa, fb = a.select(rot), b.select(rot)
            for g in range(B):
                assert a.path_len(g) == b.path_len(g), (move, c, g, a.path_len(g), b.path_len(g))
                if muzero:
                    assert a.leaf_action(g) == b.leaf_action(g) and a.path_hash(g) == b.path_hash(g), (move, c, g)
            if check_features and (not muzero or c == 0):
                assert np.array_equal(fa, fb), (move, c)
            logits = rng.normal(0.0, 2.0, size=(B, A)).astype(np.float32)
            if rng.random() < 0.25:  # quantised logits: exact prior ties, the case where std::sort's order is algorithm-defined
                logits = np.round(logits).astype(np.float32)
            e = np.exp(logits - logits.max(axis=1, keepdims=True))
            policy = (e / e.sum(axis=1, keepdims=True)).astype(np.float32)
            value = np.tanh(rng.normal(0.0, 0.7, size=B)).astype(np.float32)
            a.apply(policy, logits, value, nz)
            b.apply(policy, logits, value, nz)
        for g in range(B):
            ra, rb = a.root(g), b.root(g)
            assert ra["num_children"] == rb["num_children"], (move, g)
            k = ra["num_children"]
            assert ra["root_count"] == rb["root_count"] == S + 1 and np.float32(ra["root_mean"]) == np.float32(rb["root_mean"]), (move, g)
            assert np.array_equal(ra["action"][:k], rb["action"][:k]), (move, g)
            for name in ("count", "mean", "policy", "logit", "noise", "value"):
                assert np.array_equal(ra[name][:k].view(np.uint32), rb[name][:k].view(np.uint32)), (move, g, name)
            if gumbel:
                act = a.gumbel_best_action(g)
                assert act == b.gumbel_best_action(g), (move, g)
            else:
                act = int(ra["action"][int(np.argmax(ra["count"][:k]))]) if rng.random() < 0.5 else int(ra["action"][rng.integers(0, k)])
            assert a.play(g, act) == 1 and b.play(g, act) == 1, (move, g, act)
            ta, tb = a.root_terminal(g), b.root_terminal(g)
            assert ta == tb, (move, g)
            if ta:
                a.reset_game(g)
                b.reset_game(g)
            compared += 1
    return compared
