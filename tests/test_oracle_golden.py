"""The CPU oracle restatement (oracle/port) against vectors recorded from the compiled reference."""
import numpy as np
import pytest

import golden_replay
import oracle_lib

CASES = {
    "ttt_s50_b2": (oracle_lib.GAME_TICTACTOE, 3),
    "ttt_s50_b1_det": (oracle_lib.GAME_TICTACTOE, 3),
    "go5_s24_b2": (oracle_lib.GAME_GO, 5),
    "go9_s32_b2": (oracle_lib.GAME_GO, 9),
    "go19_s8_b2": (oracle_lib.GAME_GO, 19),
    "nogo9_s8_b2": (oracle_lib.GAME_NOGO, 9),
    "gomoku15_s8_b2": (oracle_lib.GAME_GOMOKU, 15),
    "hex11_s8_b2": (oracle_lib.GAME_HEX, 11),
    "killallgo7_s16_b2": (oracle_lib.GAME_KILLALLGO, 7),
    "go5_mz_s16_b2": (oracle_lib.GAME_GO, 5),
    "ttt_gmz_s16_b2": (oracle_lib.GAME_TICTACTOE, 3),
    "othello_gmz_s16_b2": (oracle_lib.GAME_OTHELLO, 8),
    "othello_gmz_s32_m8_b2": (oracle_lib.GAME_OTHELLO, 8),
    "othello_mz_s24_b2": (oracle_lib.GAME_OTHELLO, 8),
    "go5_gmz_s64_m12_b2": (oracle_lib.GAME_GO, 5),
    "go5_gmz_s100_m14_b2": (oracle_lib.GAME_GO, 5),
    "go9_s400_b2": (oracle_lib.GAME_GO, 9),
    "go19_s800_b2": (oracle_lib.GAME_GO, 19),
}


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference_recording(oracle, name):
    game, n = CASES[name]
    case = golden_replay.load_case(name)
    golden_replay.assert_tie_free(case)
    eng = oracle_lib.OracleSearch(oracle, game, n, int(case["B"]), int(case["S"]), **oracle_lib.conf_overrides(case["conf"]))
    checked = golden_replay.replay(eng, case)
    assert checked >= case["move_game"].size - int(case["B"])


@pytest.mark.parametrize("name", ["atari_mz_s20_b2", "atari_mz_s50_b2_det", "atari_mz_s18_gumbel_b2"])
def test_oracle_matches_reference_recording_atari(oracle, name):
    """Atari MuZero (-DATARI reference over the synthetic frame source): rewards in the tree, the value-bound multiset and its
    rescaling, the #if ATARI init-Q, one-player value sign, AtariEnv::getFeatures planes"""
    case = golden_replay.load_case(name)
    eng = oracle_lib.OracleSearch(oracle, oracle_lib.GAME_ATARI, 6, int(case["B"]), int(case["S"]), **oracle_lib.conf_overrides(case["conf"]))
    checked = golden_replay.replay_atari(eng, case)
    assert checked >= case["move_game"].size - int(case["B"])
    assert int(case["bound_size"].max()) > 2 and float(np.abs(case["child_reward"]).max()) > 0  # the recording exercises what it is here for


@pytest.mark.parametrize("name,net,dims", [
    ("ttt_s50_b2", "ttt_az_2bx32", (4, 3, 3, 32, 2, 9, 256)),
    ("go5_s24_b2", "go5_az_1bx16", (18, 5, 5, 16, 1, 26, 64)),
])
def test_oracle_net_matches_reference_outputs(oracle, name, net, dims):
    """fp32 restatement of the network vs the outputs the reference's TorchScript forward produced."""
    import os
    torch = pytest.importorskip("torch")
    path = os.path.join(oracle_lib.ROOT, "oracle", "_ref", "nets", net + ".pt")
    if not os.path.exists(path):
        pytest.skip("net fixture not generated (oracle/gen_nets.py needs /root/reference)")
    case = golden_replay.load_case(name)
    sd = torch.jit.load(path).state_dict()
    h = oracle.mzo_net_create(*dims)
    for k, v in sd.items():
        if k.endswith("num_batches_tracked"):
            continue
        a = np.ascontiguousarray(v.float().numpy())
        assert oracle.mzo_net_set(h, k.encode(), oracle_lib.fptr(a), a.size) == 0
    n, A, F = 64, int(case["A"]), int(case["F"])
    feats = np.unpackbits(case["eval_features"][:n], axis=1)[:, :F].astype(np.float32)
    pol, lg, val = np.zeros((n, A), np.float32), np.zeros((n, A), np.float32), np.zeros(n, np.float32)
    oracle.mzo_net_forward(h, oracle_lib.fptr(feats), n, oracle_lib.fptr(pol), oracle_lib.fptr(lg), oracle_lib.fptr(val))
    oracle.mzo_net_destroy(h)
    assert np.abs(lg - case["eval_logits"][:n]).max() < 1e-5
    assert np.abs(pol - case["eval_policy"][:n]).max() < 1e-6
    assert np.abs(val - case["eval_value"][:n]).max() < 1e-5


def test_oracle_gumbel_policy_matches_reference_records(oracle):
    """GumbelZero::getMCTSPolicy (gumbel_zero.cpp:9-59): the P[...] tags of the records the compiled reference printed."""
    import re
    case = golden_replay.load_case("othello_gmz_s16_b2")
    eng = oracle_lib.OracleSearch(oracle, oracle_lib.GAME_OTHELLO, 8, int(case["B"]), int(case["S"]), **oracle_lib.conf_overrides(case["conf"]))
    got = {g: [] for g in range(int(case["B"]))}

    def on_move(g, m, engine):
        a, p = engine.gumbel_policy(g)
        got[g].append((int(case["move_action"][m]), dict(zip(a.tolist(), p.tolist()))))

    golden_replay.replay(eng, case, on_move=on_move)
    lines = [str(x) for x in case["selfplay_lines"]]
    assert lines
    compared = 0
    for line in lines:
        moves = re.findall(r";[BW]\[(\d+)\]P\[([^\]]*)\]", line)
        acts = [int(a) for a, _ in moves]
        g = next(g for g in got if [a for a, _ in got[g][:len(acts)]] == acts)
        for (a, ptag), (a2, dist) in zip(moves, got[g]):
            want = {int(k): float(v) for k, v in (kv.split(":") for kv in ptag.split(","))}
            assert set(want) == set(dist), (want, dist)
            for k in want:
                assert abs(want[k] - dist[k]) <= 1e-5 * max(1.0, abs(want[k])) + 5e-7, (k, want[k], dist[k])
            compared += 1
    assert compared > 100


ENV_CASES = {"env_ttt": (oracle_lib.GAME_TICTACTOE, 3), "env_go5": (oracle_lib.GAME_GO, 5), "env_go9": (oracle_lib.GAME_GO, 9),
             "env_go9_situational": (oracle_lib.GAME_GO, 9), "env_go19": (oracle_lib.GAME_GO, 19), "env_othello8": (oracle_lib.GAME_OTHELLO, 8), "env_nogo9": (oracle_lib.GAME_NOGO, 9),
             "env_gomoku15": (oracle_lib.GAME_GOMOKU, 15), "env_gomoku15_freestyle": (oracle_lib.GAME_GOMOKU, 15),
             "env_hex11": (oracle_lib.GAME_HEX, 11), "env_hex11_noswap": (oracle_lib.GAME_HEX, 11), "env_killallgo7": (oracle_lib.GAME_KILLALLGO, 7)}


@pytest.mark.parametrize("name", list(ENV_CASES))
def test_oracle_env_matches_reference_playouts(oracle, name):
    """random legal playouts of the reference's own environments (the reference's `-mode env_test` idea): legal sets, rotated
    feature planes, terminal flags and the evaluation score after EVERY move (Tromp-Taylor on hundreds of mid-game positions)"""
    import env_replay
    game, n = ENV_CASES[name]
    case = env_replay.load(name)
    flags = (0 if "exactly_five_stones=false" in str(case["conf"]) else 1) | (2 if "outer_open" in str(case["conf"]) else 0)
    flags |= (0 if "hex_use_swap_rule=false" in str(case["conf"]) else 4)
    eng = oracle_lib.OracleSearch(oracle, game, n, 1, 1, ko_situational=int("situational" in str(case["conf"])), gomoku_flags=flags)
    state = {"i": 0}

    def score(e):
        return oracle.mzo_env_eval_score(oracle.mzo_root_env(e.h, 0), 0)

    assert env_replay.replay(eng, case, check_score=score) == case["game"].size
    # the score of every intermediate position as well
    eng.reset_game(0)
    for i in range(case["game"].size):
        if case["step"][i] == 0:
            eng.reset_game(0)
        assert eng.play(0, int(case["action"][i])) == 1
        assert score(eng) == case["score_after"][i], i


THINK_CASES = {"think_ttt_s50_k4": (oracle_lib.GAME_TICTACTOE, 3), "think_go5_s60_k8": (oracle_lib.GAME_GO, 5), "think_go9_s100_k16_det": (oracle_lib.GAME_GO, 9),
               "think_go5_s23_k5": (oracle_lib.GAME_GO, 5), "think_othello_mz_s30_k6": (oracle_lib.GAME_OTHELLO, 8), "think_go5_mz_s20_k4": (oracle_lib.GAME_GO, 5)}


@pytest.mark.parametrize("name", list(THINK_CASES))
def test_oracle_think_matches_reference_recording(oracle, name):
    """console search (ZeroActor::think, actor_mcts_think_batch_size > 1): selection under virtual loss, duplicate leaves, short last batch"""
    game, n = THINK_CASES[name]
    case = golden_replay.load_case(name)
    eng = oracle_lib.OracleSearch(oracle, game, n, 1, int(case["S"]), **oracle_lib.conf_overrides(case["conf"]))
    assert golden_replay.replay_think(eng, case) == case["move_action"].size
