"""Replays an environment differential-playout fixture (tests/golden/env_*.npz, recorded from the reference's own environment
classes by oracle/gen_env_golden.py) through a search engine with one game and one simulation: before every move the root
position's feature planes under the recorded rotation and its legal action set (= the root's children after one evaluation) must
match; after it the terminal flag, and at the end of a game the score."""
import numpy as np

import golden_replay


def replay(engine, case, check_score=None):
    A, F = int(case["A"]), int(case["F"])
    uniform = np.full((1, A), 1.0 / A, np.float32)
    zeros, value = np.zeros((1, A), np.float32), np.zeros(1, np.float32)
    steps = 0
    for i in range(case["game"].size):
        if case["step"][i] == 0:
            engine.reset_game(0)
        feats = engine.select(np.array([case["rotation"][i]], np.uint8))
        want = np.unpackbits(case["features"][i])[:F].astype(np.float32)
        assert np.array_equal(feats[0], want), f"feature planes differ at record {i}"
        engine.apply(uniform, zeros, value, None)
        r = engine.root(0)
        legal = np.nonzero(np.unpackbits(case["legal"][i])[:A])[0]
        assert r["num_children"] == legal.size and np.array_equal(np.sort(r["action"][:legal.size]), legal), f"legal set differs at record {i}"
        assert engine.play(0, int(case["action"][i])) == 1, f"legal move refused at record {i}"
        assert engine.root_terminal(0) == bool(case["terminal_after"][i]), f"terminal flag differs at record {i}"
        if check_score is not None and case["terminal_after"][i]:
            assert check_score(engine) == case["score_after"][i], f"final score differs at record {i}"
        steps += 1
    return steps


def load(name):
    return golden_replay.load_case(name)
