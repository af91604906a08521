"""Host side of the learner data path (minizero_b200/learner_data.py), CPU only: record parsing and policy / value targets against
records the compiled reference printed; the rotation table against the oracle's restatement of utils/rotation.h."""
import numpy as np

import golden_replay
import oracle_lib
from minizero_b200 import learner_data


def test_rotate_action_matches_rotation_h(oracle):
    for n in (3, 5, 9, 19):
        for r in range(8):
            for a in range(n * n + 1):
                assert learner_data.rotate_action(a, r, n) == oracle.mzo_rotate_position(r, a, n), (n, r, a)


def test_records_parse_and_targets_follow_the_reference_loader():
    z = golden_replay.load_case("go5_s24_b2")
    lines = [str(l) for l in z["selfplay_lines"]]
    assert lines
    for line in lines:
        rec = learner_data.parse_record(line)
        game_len = int(line.split(" ")[3])
        assert len(rec) == game_len and rec.tags["GM"] == "go_5x5" and rec.tags["SZ"] == "5"
        assert rec.data_range() == (0, game_len - 1)
        assert rec.players[:4] == [1, 2, 1, 2][:game_len]
        assert rec.value(0) == np.float32(float(line.split(" ")[4]))
        for pos in (0, game_len // 2, game_len - 1):
            for rot in (0, 3, 5):
                p = rec.policy(pos, rot, 26, 5)
                assert abs(float(p.sum()) - 1.0) < 1e-5 and p.min() >= 0
                counts = dict((int(a), float(c)) for a, c in (t.split(":") for t in rec.infos[pos]["P"].split(",")))
                total = sum(counts.values())
                for a, c in counts.items():
                    assert abs(float(p[learner_data.rotate_action(a, rot, 5)]) - c / total) < 1e-6
        assert np.allclose(rec.policy(game_len + 2, 0, 26, 5), 1.0 / 26)  # absorbing state
    # the intermediate sequences of long games carry their DLEN window
    z = golden_replay.load_case("go5_seq_s8_b2")
    rec = learner_data.parse_record(str(z["selfplay_lines"][0]))
    lo, hi = rec.data_range()
    assert hi - lo + 1 == int(str(z["selfplay_lines"][0]).split(" ")[2])


def test_muzero_targets_follow_the_reference_loader():
    """setMuZeroTrainingData's per-step targets (data_loader.cpp:159-200) for a board-game record: action planes (rotated one-hot, empty for a
    pass, the reference's random plane past the end), policies, the game's return as every value, R tags as rewards"""

    class StubEngine:  # only what muzero_batch reads besides the device call, which is replaced below
        game, A, board_size = 1, 26, 5

    z = golden_replay.load_case("go5_mz_s16_b2")
    recs = [learner_data.parse_record(str(l)) for l in z["selfplay_lines"]]
    assert recs
    rec = recs[0]
    L = len(rec)
    draws = iter([7, 3, 25, 11, 2, 19, 5, 23])
    orig = learner_data.alphazero_batch
    learner_data.alphazero_batch = lambda e, r, p: (np.zeros((len(p), 18 * 25), np.float32), None, None)
    try:
        picks = [(0, 0, 0), (0, L - 2, 3), (0, L // 2, 6)]
        feats, act, pol, val, rew = learner_data.muzero_batch(StubEngine(), recs, picks, 5, lambda: next(draws))
    finally:
        learner_data.alphazero_batch = orig
    assert act.shape == (3, 5, 25) and pol.shape == (3, 6, 26) and val.shape == (3, 6) and rew.shape == (3, 5)
    assert np.all(val == np.float32(float(rec.tags["RE"])))
    for j, (_, p, q) in enumerate(picks):
        for step in range(5):
            pos = p + step
            if pos < L:
                a = rec.actions[pos]
                want = np.zeros(25, np.float32)
                if a != 25:
                    want[learner_data.rotate_action(a, q, 5)] = 1.0
                assert np.array_equal(act[j, step], want)
                assert rew[j, step] == np.float32(float(rec.infos[pos]["R"]))
            else:
                assert act[j, step].sum() in (0.0, 1.0) and rew[j, step] == 0.0
        for step in range(6):
            assert abs(float(pol[j, step].sum()) - 1.0) < 1e-5
            if p + step >= L:
                assert np.allclose(pol[j, step], 1.0 / 26)
    # positions past the end consumed the generator, in the reference's order: pick 1 runs 4 steps past the end of the game (L - 2 + 2 .. + 4, action planes only for < K)
    assert next(draws, None) is not None
