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
