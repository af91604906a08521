"""Learner data path on the device: BaseEnvLoader::getFeatures — replay a record to a position and emit the
rotated feature planes — for a batch of samples, against planes the reference's own environment classes produced. -m gpu."""
import numpy as np
import pytest

import env_replay

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,game,n", [("env_ttt", 0, 3), ("env_go5", 1, 5), ("env_go9", 1, 9), ("env_go19", 1, 19), ("env_othello8", 2, 8), ("env_nogo9", 3, 9),
                                         ("env_gomoku15", 4, 15), ("env_hex11", 5, 11)])
def test_replay_features_match_reference_environment(name, game, n):
    """every recorded position of the reference's random playouts (captures, ko, passes, flips, swaps) is rebuilt from its game's action
    list alone, in batches, under the recorded rotation: the planes must equal what the reference's Environment::getFeatures returned"""
    import minizero_b200
    case = env_replay.load(name)
    A, F = int(case["A"]), int(case["F"])
    games = {}
    for i in range(case["game"].size):
        games.setdefault(int(case["game"][i]), []).append(i)
    max_len = max(len(v) for v in games.values())
    samples = [(g, k, idx) for g, recs in games.items() for k, idx in enumerate(recs)]  # position k of game g was recorded at index idx
    batch = 64
    eng = minizero_b200.Engine(game, n, batch, 1)
    checked = 0
    for s0 in range(0, len(samples), batch):
        chunk = samples[s0:s0 + batch]
        actions = np.full((len(chunk), max_len), -1, np.int32)
        for j, (g, k, idx) in enumerate(chunk):
            recs = games[g]
            actions[j, :len(recs)] = case["action"][recs]
        pos = np.array([k for _, k, _ in chunk], np.int32)
        rot = np.array([case["rotation"][idx] for _, _, idx in chunk], np.uint8)
        feats = eng.replay_features(actions, pos, rot)
        for j, (g, k, idx) in enumerate(chunk):
            want = np.unpackbits(case["features"][idx])[:F].astype(np.float32)
            assert np.array_equal(feats[j], want), (name, g, k)
            checked += 1
    assert checked == case["game"].size
    eng.close()
