"""Parity of the CUDA path (through the C ABI, libmzb200.so) with the reference. Needs a B200: -m gpu."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

import golden_replay
import oracle_lib

pytestmark = pytest.mark.gpu

ROOT = oracle_lib.ROOT
NETS = os.path.join(ROOT, "oracle", "_ref", "nets")
CASES = {"ttt_s50_b2": (0, 3), "ttt_s50_b1_det": (0, 3), "go5_s24_b2": (1, 5), "go9_s32_b2": (1, 9), "go19_s8_b2": (1, 19), "nogo9_s8_b2": (3, 9), "gomoku15_s8_b2": (4, 15), "hex11_s8_b2": (5, 11), "killallgo7_s16_b2": (7, 7),
         "go5_mz_s16_b2": (1, 5), "ttt_gmz_s16_b2": (0, 3), "othello_gmz_s16_b2": (2, 8), "othello_gmz_s32_m8_b2": (2, 8), "othello_mz_s24_b2": (2, 8),
         # odd Gumbel sample sizes and BASELINE search lengths: 400 simulations with the 6b x 256 net, 19x19 with 800 simulations
         "go5_gmz_s64_m12_b2": (1, 5), "go5_gmz_s100_m14_b2": (1, 5), "go9_s400_b2": (1, 9), "go19_s800_b2": (1, 19)}


def engine(*args, **kw):
    import minizero_b200
    return minizero_b200.Engine(*args, **kw)


@pytest.mark.parametrize("name", list(CASES))
def test_tree_and_env_kernels_replay_reference_recording(name):
    """bit-exact: feature planes of every leaf, path lengths, root child tables (visit counts, means, priors)"""
    game, n = CASES[name]
    case = golden_replay.load_case(name)
    golden_replay.assert_tie_free(case)
    eng = engine(game, n, int(case["B"]), int(case["S"]), **oracle_lib.conf_overrides(case["conf"]))
    checked = golden_replay.replay(eng, case)
    assert checked >= case["move_game"].size - int(case["B"])
    eng.close()


def torchscript(net):
    torch = pytest.importorskip("torch")
    path = os.path.join(NETS, net + ".pt")
    if not os.path.exists(path):
        pytest.skip("net fixture missing (oracle/gen_nets.py needs the reference checkout)")
    return torch, torch.jit.load(path, map_location="cpu").eval(), path


@pytest.mark.parametrize("net,game,n,batch", [("ttt_az_2bx32", 0, 3, 64), ("go5_az_1bx16", 1, 5, 32), ("go9_az_1bx16", 1, 9, 32), ("go9_az_2bx64", 1, 9, 32),
                                              ("go9_az_6bx256", 1, 9, 256)])
def test_network_matches_torchscript_fp32(net, game, n, batch):
    """policy / value logits within 1e-3 (north_star tolerance) of the reference TorchScript module run in fp32 on the CPU"""
    torch, m, path = torchscript(net)
    eng = engine(game, n, batch, 8)
    eng.load_network(path)
    rng = np.random.default_rng(7)
    C = m.get_num_input_channels()
    feats = (rng.random((batch, C, n, n)) < 0.25).astype(np.float32)
    feats[0] = 0.0
    feats[1] = 1.0
    with torch.no_grad():
        ref = m(torch.from_numpy(feats))
    pol, lg, val = eng.eval_batch(feats)
    tol = 1e-3
    assert np.abs(lg - ref["policy_logit"].numpy()).max() < tol
    assert np.abs(val - ref["value"].numpy().reshape(-1)).max() < tol
    assert np.abs(pol - ref["policy"].numpy()).max() < tol
    eng.close()


def test_network_matches_oracle_port_random_weights():
    import __graft_entry__ as ge
    rng = np.random.default_rng(3)
    dims = dict(num_input_channels=18, input_height=5, input_width=5, num_hidden_channels=48, num_blocks=2, action_size=26, num_value_hidden_channels=32,
                discrete_value_size=1)
    state = ge.make_random_state(dims, rng)
    net = oracle_lib.OracleNet(oracle_lib.load(), dims, state)
    eng = engine(1, 5, 16, 8)
    eng.load_network((dims, state))
    feats = (rng.random((16, 18, 5, 5)) < 0.3).astype(np.float32)
    pol, lg, val = eng.eval_batch(feats)
    p2, l2, v2 = net.forward(feats.reshape(16, -1))
    assert np.abs(lg - l2).max() < 1e-3 and np.abs(val - v2).max() < 1e-3 and np.abs(pol - p2).max() < 1e-3
    eng.close()


def run_search_vs_oracle(game, n, B, S, net_path, moves, seed):
    """whole-move on-device searches (CUDA graph: tree kernels + tensor-core network) against the oracle fed with the
    engine's own network outputs: visit counts must agree exactly, move after move, with rotations and Dirichlet noise"""
    lib = oracle_lib.load()
    eng = engine(game, n, B, S)
    eng.load_network(net_path)
    ev = engine(game, n, B, 2)  # second engine: network-only, so the searching engine's buffers stay untouched
    ev.load_network(net_path)
    orc = oracle_lib.OracleSearch(lib, game, n, B, S)
    rng = np.random.default_rng(seed)
    A = eng.A
    for move in range(moves):
        rot = rng.integers(0, 8, size=(S + 1, B)).astype(np.uint8)
        noise = rng.dirichlet([0.3] * A, size=B).astype(np.float32)
        eng.set_search_inputs(rot, noise)
        eng.search()
        for c in range(S + 1):
            feats = orc.select(rot[c])
            pol, lg, val = ev.eval_batch(feats)
            orc.apply(pol, lg, val, noise)
        actions = np.zeros(B, np.int32)
        for g in range(B):
            a, b = eng.root(g), orc.root(g)
            assert a["num_children"] == b["num_children"], (move, g)
            k = a["num_children"]
            assert np.array_equal(a["action"][:k], b["action"][:k]), (move, g)
            assert np.array_equal(a["count"][:k], b["c_count"][:k] if "c_count" in b else b["count"][:k]), (move, g)
            assert np.array_equal(a["mean"][:k].view(np.uint32), b["mean"][:k].view(np.uint32)), (move, g)
            assert a["root_count"] == S + 1 and a["count"][:k].sum() == S
            actions[g] = a["action"][rng.integers(0, k)]  # any legal move keeps both sides in step
        res = eng.play_all(actions)
        for g in range(B):
            assert res["applied"][g] == 1 and orc.play(g, int(actions[g])) == 1
            assert bool(res["terminal"][g]) == orc.root_terminal(g)
            if res["terminal"][g]:
                eng.reset_game(g)
                orc.reset_game(g)
    eng.close()
    ev.close()


def test_on_device_search_matches_oracle_tictactoe():
    torch, m, path = torchscript("ttt_az_2bx32")
    run_search_vs_oracle(0, 3, 8, 50, path, moves=12, seed=1)


def test_on_device_search_matches_oracle_go5():
    torch, m, path = torchscript("go5_az_1bx16")
    run_search_vs_oracle(1, 5, 8, 24, path, moves=60, seed=2)


def test_on_device_search_matches_oracle_go9():
    torch, m, path = torchscript("go9_az_2bx64")
    run_search_vs_oracle(1, 9, 8, 32, path, moves=20, seed=3)


def test_on_device_search_matches_oracle_nogo9():
    """NoGo (GoEnv with its own legality, end and result): random moves carry the games to their end several times"""
    torch, m, path = torchscript("nogo9_az_1bx16")
    run_search_vs_oracle(3, 9, 4, 16, path, moves=90, seed=9)


def test_on_device_search_matches_oracle_go9_deep_paths():
    """240 simulations with a random 1-block net: the tree degenerates into chains deeper than 48 levels, which the device
    re-evaluates one THREAD per level (visited-list and scan forms) instead of one warp per level"""
    torch, m, path = torchscript("go9_az_1bx16")
    run_search_vs_oracle(1, 9, 4, 240, path, moves=3, seed=8)


def test_live_reference_stepper_replay_go9():
    """run the compiled reference itself (oracle/_ref, shipped with the snapshot) on this box and replay it"""
    binary = os.path.join(ROOT, "oracle", "_ref", "ref_stepper_go")
    net = os.path.join(NETS, "go9_az_1bx16.pt")
    if not (os.path.exists(binary) and os.path.exists(net)):
        pytest.skip("oracle/_ref not built")
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import gen_golden
    conf = ("env_board_size=9:actor_num_simulation=64:zero_num_parallel_games=4:zero_num_threads=1:program_seed=11:program_auto_seed=false:"
            "program_quiet=true:nn_type_name=alphazero:nn_file_name=" + net)
    with tempfile.TemporaryDirectory() as d:
        subprocess.run([binary, conf, d, "80"], check=True, capture_output=True, timeout=600)
        meta = dict(line.split() for line in open(os.path.join(d, "meta.txt")))
        case = gen_golden.read_case(d, int(meta["A"]), int(meta["F"]))
        case.update(A=int(meta["A"]), F=int(meta["F"]), S=int(meta["S"]), B=int(meta["B"]))
    golden_replay.assert_tie_free(case)
    eng = engine(1, 9, 4, 64)
    assert golden_replay.replay(eng, case) >= 76
    eng.close()


def test_full_size_search_invariants():
    """BASELINE config 2 at full size (256 games, 400 simulations, 6b x 256): size-independent properties"""
    torch, m, path = torchscript("go9_az_6bx256")
    B, S = 256, 400
    eng = engine(1, 9, B, S)
    eng.load_network(path)
    rng = np.random.default_rng(5)
    rot = rng.integers(0, 8, size=(S + 1, B)).astype(np.uint8)
    noise = rng.dirichlet([0.03] * eng.A, size=B).astype(np.float32)
    eng.set_search_inputs(rot, noise)
    ms = eng.search()
    r = eng.get_roots()
    assert np.all(r["root_count"] == S + 1)
    assert np.all(r["num_children"] == 82)  # empty board: every point and the pass are legal
    assert np.all(r["count"].sum(axis=1) == S)
    assert np.all(np.sort(r["action"], axis=1) == np.arange(82))
    assert np.all(np.abs(r["root_mean"]) <= 1.0) and np.all(np.isfinite(r["mean"]))
    # idempotence: the same inputs from the same position give the same tree
    eng.reset_game(-1)
    eng.set_search_inputs(rot, noise)
    eng.search()
    r2 = eng.get_roots()
    assert np.array_equal(r["count"], r2["count"]) and np.array_equal(r["mean"].view(np.uint32), r2["mean"].view(np.uint32))
    print("full-size search: %.1f ms, %.0f evals/s" % (ms, B * (S + 1) / ms * 1e3))
    eng.close()


# ---- MuZero / Gumbel (BASELINE configs[2]: Othello 8x8 Gumbel MuZero) ------------------------------------------------

@pytest.mark.parametrize("net,batch,game,n", [("othello_mz_1bx32", 32, 2, 8), ("othello_mz_3bx128", 64, 2, 8), ("go5_mz_1bx16", 16, 1, 5), ("ttt_mz_1bx16", 16, 0, 3)])
def test_muzero_network_matches_torchscript_fp32(net, batch, game, n):
    """initial_inference and recurrent_inference (network/py/muzero_network.py:136-150) against the reference TorchScript
    module in fp32 on the CPU: logits / value / policy within 1e-3; the scaled hidden state (values in [0, 1], kept in fp16
    here) within 2e-3"""
    torch, m, path = torchscript(net)
    eng = engine(game, n, batch, 8, muzero=1)
    eng.load_network(path)
    rng = np.random.default_rng(11)
    C = m.get_num_input_channels()
    feats = (rng.random((batch, C, n, n)) < 0.3).astype(np.float32)
    turn = rng.integers(0, 2, size=batch)
    feats[:, C - 2] = (turn == 0)[:, None, None]
    feats[:, C - 1] = (turn == 1)[:, None, None]
    with torch.no_grad():
        ref = m.initial_inference(torch.from_numpy(feats))
    pol, lg, val, hid = eng.eval_initial(feats)
    assert np.abs(lg - ref["policy_logit"].numpy()).max() < 1e-3
    assert np.abs(val - ref["value"].numpy().reshape(-1)).max() < 1e-3
    assert np.abs(pol - ref["policy"].numpy()).max() < 1e-3
    ref_hid = ref["hidden_state"].numpy().reshape(batch, -1)
    assert hid.min() >= 0.0 and hid.max() <= 1.0 and np.abs(hid - ref_hid).max() < 2e-3
    # recurrent inference from the REFERENCE's hidden states; actions include the pass (all-zero plane, othello.cpp:257-262)
    actions = rng.integers(0, eng.A, size=batch).astype(np.int32)
    if eng.A > n * n:
        actions[:2] = n * n
    planes = np.zeros((batch, 1, n * n), np.float32)
    for g in range(batch):
        if actions[g] < n * n:
            planes[g, 0, actions[g]] = 1.0
    with torch.no_grad():
        ref2 = m.recurrent_inference(ref["hidden_state"], torch.from_numpy(planes.reshape(batch, 1, n, n)))
    pol2, lg2, val2, hid2 = eng.eval_recurrent(ref_hid, actions)
    assert np.abs(lg2 - ref2["policy_logit"].numpy()).max() < 1e-3
    assert np.abs(val2 - ref2["value"].numpy().reshape(-1)).max() < 1e-3
    assert np.abs(pol2 - ref2["policy"].numpy()).max() < 1e-3
    assert np.abs(hid2 - ref2["hidden_state"].numpy().reshape(batch, -1)).max() < 2e-3
    eng.close()


def run_muzero_search_vs_oracle(B, S, net_path, moves, seed, game=2, n=8, **opts):
    """whole-move on-device MuZero searches (CUDA graph: tree kernels, hidden-state gather, representation / dynamics towers,
    hidden-state scaling, heads) against the oracle fed with the engine's own network outputs, move after move"""
    lib = oracle_lib.load()
    eng = engine(game, n, B, S, muzero=1, **opts)
    eng.load_network(net_path)
    ev = engine(game, n, B, 2, muzero=1)  # network-only engine
    ev.load_network(net_path)
    orc = oracle_lib.OracleSearch(lib, game, n, B, S, muzero=1, **opts)
    rng = np.random.default_rng(seed)
    A = eng.A
    gumbel = bool(opts.get("use_gumbel"))
    for move in range(moves):
        noise = (rng.gumbel(size=(B, A)) if opts.get("gumbel_noise") else rng.dirichlet([0.3] * A, size=B)).astype(np.float32)
        eng.set_search_inputs(None, noise)
        eng.search()
        store = None
        for c in range(S + 1):
            feats = orc.select(None)
            lens = np.array([orc.path_len(g) for g in range(B)])
            if c == 0:
                assert np.all(lens == 1)
                pol, lg, val, hid = ev.eval_initial(feats)
                store = np.zeros((B, S + 1) + hid.shape[1:], np.float32)
            else:
                assert np.all(lens > 1)
                parent = np.array([orc.leaf_parent_slot(g) for g in range(B)])
                acts = np.array([orc.leaf_action(g) for g in range(B)], np.int32)
                pol, lg, val, hid = ev.eval_recurrent(store[np.arange(B), parent], acts)
            store[:, c] = hid
            orc.apply(pol, lg, val, noise)
        best = eng.gumbel_best_actions() if gumbel else None
        actions = np.zeros(B, np.int32)
        for g in range(B):
            a, b = eng.root(g), orc.root(g)
            assert a["num_children"] == b["num_children"], (move, g)
            k = a["num_children"]
            assert np.array_equal(a["action"][:k], b["action"][:k]), (move, g)
            assert np.array_equal(a["count"][:k], b["count"][:k]), (move, g, a["count"][:k], b["count"][:k])
            assert np.array_equal(a["mean"][:k].view(np.uint32), b["mean"][:k].view(np.uint32)), (move, g)
            assert np.array_equal(a["logit"][:k].view(np.uint32), b["logit"][:k].view(np.uint32)), (move, g)
            assert a["root_count"] == S + 1 and a["count"][:k].sum() == S
            if gumbel:
                assert best[g] == orc.gumbel_best_action(g), (move, g)
                actions[g] = best[g]
            else:
                actions[g] = a["action"][int(np.argmax(a["count"][:k]))]
        res = eng.play_all(actions)
        for g in range(B):
            assert res["applied"][g] == 1 and orc.play(g, int(actions[g])) == 1
            assert bool(res["terminal"][g]) == orc.root_terminal(g)
            if res["terminal"][g]:
                eng.reset_game(g)
                orc.reset_game(g)
    eng.close()
    ev.close()


def test_on_device_gumbel_muzero_search_matches_oracle_othello():
    torch, m, path = torchscript("othello_mz_1bx32")
    run_muzero_search_vs_oracle(8, 16, path, moves=66, seed=4, use_gumbel=1, gumbel_noise=1, gumbel_sample_size=16)


def test_on_device_gumbel_muzero_search_with_halving_matches_oracle_othello():
    torch, m, path = torchscript("othello_mz_3bx128")
    run_muzero_search_vs_oracle(8, 32, path, moves=12, seed=5, use_gumbel=1, gumbel_noise=1, gumbel_sample_size=8)


def test_on_device_puct_muzero_search_matches_oracle_othello():
    torch, m, path = torchscript("othello_mz_1bx32")
    run_muzero_search_vs_oracle(8, 24, path, moves=10, seed=6)


def test_on_device_muzero_search_matches_oracle_go5():
    torch, m, path = torchscript("go5_mz_1bx16")
    run_muzero_search_vs_oracle(8, 16, path, moves=30, seed=7, game=1, n=5)


def test_on_device_gumbel_muzero_search_matches_oracle_tictactoe():
    torch, m, path = torchscript("ttt_mz_1bx16")
    run_muzero_search_vs_oracle(8, 16, path, moves=9, seed=8, game=0, n=3, use_gumbel=1, gumbel_noise=1, gumbel_sample_size=4)


def test_full_size_gumbel_muzero_invariants():
    """BASELINE config 3 at full size (512 games, Gumbel n=16 m=16, 3b x 128 MuZero): size-independent properties"""
    torch, m, path = torchscript("othello_mz_3bx128")
    B, S = 512, 16
    eng = engine(2, 8, B, S, muzero=1, use_gumbel=1, gumbel_noise=1, gumbel_sample_size=16)
    eng.load_network(path)
    rng = np.random.default_rng(9)
    noise = rng.gumbel(size=(B, eng.A)).astype(np.float32)
    eng.set_search_inputs(None, noise)
    ms = eng.search()
    r = eng.get_roots()
    assert np.all(r["root_count"] == S + 1)
    assert np.all(r["num_children"] == 4)  # the four opening moves of Othello
    assert np.all(np.sort(r["action"][:, :4], axis=1) == np.array([20, 29, 34, 43]))
    assert np.all(r["count"][:, :4] == 4)  # budget 1 per candidate, no halving (gumbel_zero.cpp:99,109): 16 simulations over 4 candidates
    assert np.all(np.isfinite(r["mean"])) and np.all(np.abs(r["root_mean"]) <= 1.0)
    eng.reset_game(-1)
    eng.set_search_inputs(None, noise)
    eng.search()
    r2 = eng.get_roots()
    assert np.array_equal(r["count"], r2["count"]) and np.array_equal(r["mean"].view(np.uint32), r2["mean"].view(np.uint32))
    print("config-3 search: %.2f ms, %.0f evals/s" % (ms, B * (S + 1) / ms * 1e3))
    eng.close()


# ---- 19x19 (BASELINE configs[3]: Go 19x19 AlphaZero, 800 simulations, 128 games, 20b x 256) ------------------------------

def test_network_19x19_matches_oracle_port_random_weights():
    import __graft_entry__ as ge
    rng = np.random.default_rng(13)
    dims = dict(num_input_channels=18, input_height=19, input_width=19, num_hidden_channels=128, num_blocks=2, action_size=362, num_value_hidden_channels=64,
                discrete_value_size=1)
    state = ge.make_random_state(dims, rng)
    net = oracle_lib.OracleNet(oracle_lib.load(), dims, state)
    eng = engine(1, 19, 8, 4)
    eng.load_network((dims, state))
    assert eng.conv_layers_per_launch() == 5  # the fused tower also covers the 19x19 resident block (fewer weight stages)
    feats = (rng.random((8, 18, 19, 19)) < 0.3).astype(np.float32)
    pol, lg, val = eng.eval_batch(feats)
    p2, l2, v2 = net.forward(feats.reshape(8, -1))
    assert np.abs(lg - l2).max() < 1e-3 and np.abs(val - v2).max() < 1e-3 and np.abs(pol - p2).max() < 1e-3
    eng.close()


def test_full_size_19x19_search_invariants():
    """BASELINE config 4 at full size (128 games, 800 simulations, 20b x 256 random-init): size-independent properties"""
    import __graft_entry__ as ge
    B, S = 128, 800
    dims = dict(num_input_channels=18, input_height=19, input_width=19, num_hidden_channels=256, num_blocks=20, action_size=362, num_value_hidden_channels=256,
                discrete_value_size=1)
    eng = engine(1, 19, B, S)
    eng.load_network((dims, ge.make_random_state(dims, np.random.default_rng(17))))
    assert eng.conv_layers_per_launch() == 41
    rng = np.random.default_rng(5)
    rot = rng.integers(0, 8, size=(S + 1, B)).astype(np.uint8)
    noise = rng.dirichlet([0.03] * eng.A, size=B).astype(np.float32)
    eng.set_search_inputs(rot, noise)
    ms = eng.search()
    r = eng.get_roots()
    assert np.all(r["root_count"] == S + 1)
    assert np.all(r["num_children"] == 362)
    assert np.all(r["count"].sum(axis=1) == S)
    assert np.all(np.sort(r["action"], axis=1) == np.arange(362))
    assert np.all(np.abs(r["root_mean"]) <= 1.0) and np.all(np.isfinite(r["mean"]))
    eng.reset_game(-1)
    eng.set_search_inputs(rot, noise)
    eng.search()
    r2 = eng.get_roots()
    assert np.array_equal(r["count"], r2["count"]) and np.array_equal(r["mean"].view(np.uint32), r2["mean"].view(np.uint32))
    print("config-4 search: %.1f ms, %.0f evals/s, %.1f TFLOP/s algorithmic" % (ms, B * (S + 1) / ms * 1e3, B * (S + 1) * 17.07e9 / ms * 1e3 / 1e12))
    eng.close()


@pytest.mark.parametrize("name,game,n", [("env_ttt", 0, 3), ("env_go5", 1, 5), ("env_go9", 1, 9), ("env_go9_situational", 1, 9), ("env_go19", 1, 19),
                                         ("env_othello8", 2, 8), ("env_nogo9", 3, 9), ("env_gomoku15", 4, 15),
                                         ("env_gomoku15_freestyle", 4, 15), ("env_hex11", 5, 11), ("env_hex11_noswap", 5, 11), ("env_killallgo7", 7, 7)])
def test_device_env_matches_reference_playouts(name, game, n):
    """the device rule / feature kernels against random playouts of the reference's own environments (captures, ko and superko,
    suicide, passes, Othello flips and forced passes): legal sets, rotated planes, terminal flags, final scores"""
    import env_replay
    case = env_replay.load(name)
    eng = engine(game, n, 1, 1, ko_situational="situational" in str(case["conf"]), gomoku_exactly_five="exactly_five_stones=false" not in str(case["conf"]),
                 gomoku_outer_open="outer_open" in str(case["conf"]), hex_swap_rule="hex_use_swap_rule=false" not in str(case["conf"]))
    assert env_replay.replay(eng, case, check_score=lambda e: float(e.last_play["eval_score"][0])) == case["game"].size
    eng.close()


@pytest.mark.parametrize("name,game,board,A,sims,games,moves,opts", [
    ("go5_az", 1, 5, 26, 200, 4, 30, {}),
    ("go9_az", 1, 9, 82, 120, 4, 16, {}),
    ("nogo9_az", 3, 9, 82, 40, 4, 80, {}),
    ("hex11_az", 5, 11, 121, 24, 4, 125, {}),
    ("othello_gumbel_muzero_m8", 2, 8, 65, 32, 4, 66, dict(muzero=1, use_gumbel=1, gumbel_noise=1, gumbel_sample_size=8)),
    ("othello_muzero", 2, 8, 65, 64, 4, 20, dict(muzero=1)),
], ids=["go5_az", "go9_az", "nogo9_az", "hex11_az", "othello_gumbel_muzero_m8", "othello_muzero"])
def test_device_and_oracle_agree_on_synthetic_searches(name, game, board, A, sims, games, moves, opts):
    """differential fuzzing through the per-phase hooks of the C ABI: random priors (with exact ties), values, noise and rotations"""
    import zlib

    import fuzz_differential
    orc = oracle_lib.OracleSearch(oracle_lib.load(), game, board, games, sims, **opts)
    eng = engine(game, board, games, sims, **opts)
    n = fuzz_differential.run(eng, orc, num_actions=A, sims=sims, games=games, moves=moves, seed=zlib.crc32(name.encode()) % 1000,
                              noise=("gumbel" if opts.get("gumbel_noise") else "dirichlet"), muzero=bool(opts.get("muzero")), gumbel=bool(opts.get("use_gumbel")))
    assert n == games * moves
    eng.close()
