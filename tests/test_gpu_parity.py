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
CASES = {"ttt_s50_b2": (0, 3), "ttt_s50_b1_det": (0, 3), "go5_s24_b2": (1, 5), "go9_s32_b2": (1, 9)}


def engine(*args, **kw):
    import minizero_b200
    return minizero_b200.Engine(*args, **kw)


@pytest.mark.parametrize("name", list(CASES))
def test_tree_and_env_kernels_replay_reference_recording(name):
    """bit-exact: feature planes of every leaf, path lengths, root child tables (visit counts, means, priors)"""
    game, n = CASES[name]
    case = golden_replay.load_case(name)
    golden_replay.assert_tie_free(case)
    eng = engine(game, n, int(case["B"]), int(case["S"]))
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
