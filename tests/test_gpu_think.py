"""Console search on the device (ZeroActor::think with actor_mcts_think_batch_size = K > 1, zero_actor.cpp:36-49,129-157) through the C ABI
(mz_config.think_batch_size): recordings of the compiled reference, and whole searches against the CPU oracle. Needs a B200: -m gpu."""
import os

import numpy as np
import pytest

import golden_replay
import oracle_lib

pytestmark = pytest.mark.gpu

NETS = os.path.join(oracle_lib.ROOT, "oracle", "_ref", "nets")
CASES = {"think_ttt_s50_k4": (0, 3), "think_go5_s60_k8": (1, 5), "think_go9_s100_k16_det": (1, 9), "think_go5_s23_k5": (1, 5),
         "think_othello_mz_s30_k6": (2, 8), "think_go5_mz_s20_k4": (1, 5)}


def engine(*args, **kw):
    import minizero_b200
    return minizero_b200.Engine(*args, **kw)


@pytest.mark.parametrize("name", list(CASES))
def test_think_kernels_replay_reference_recording(name):
    """bit-exact: which lanes of every step exist, which leaves are duplicates, path lengths, planes of the evaluated leaves, root tables"""
    game, n = CASES[name]
    case = golden_replay.load_case(name)
    eng = engine(game, n, 1, int(case["S"]), think_batch_size=int(case["K"]), **oracle_lib.conf_overrides(case["conf"]))
    assert golden_replay.replay_think(eng, case, features_of_duplicates=False) == case["move_action"].size
    eng.close()


@pytest.mark.parametrize("net,game,n,trees,S,K,moves", [("go9_az_2bx64", 1, 9, 3, 64, 8, 4), ("ttt_az_2bx32", 0, 3, 2, 50, 4, 5), ("go5_az_1bx16", 1, 5, 4, 37, 6, 6)])
def test_on_device_think_search_matches_oracle(net, game, n, trees, S, K, moves):
    """whole think() searches on the device (K selections per tree and step, tower + heads on trees x K positions, host-driven step loop) against the
    oracle stepping the same trees with the network outputs of a second, network-only engine; rotations and Dirichlet noise on"""
    path = os.path.join(NETS, net + ".pt")
    if not os.path.exists(path):
        pytest.skip("net fixture missing")
    lib = oracle_lib.load()
    eng = engine(game, n, trees, S, think_batch_size=K)
    eng.load_network(path)
    ev = engine(game, n, trees * K, 2)
    ev.load_network(path)
    orc = oracle_lib.OracleSearch(lib, game, n, trees, S)
    rng = np.random.default_rng(7)
    A = eng.A
    for move in range(moves):
        rot = rng.integers(0, 8, size=(S + 1, K, trees)).astype(np.uint8)
        noise = rng.dirichlet([0.3] * A, size=trees).astype(np.float32)
        full_noise = np.zeros((eng.B, A), np.float32)
        full_noise[:trees] = noise
        eng.set_search_inputs(rot, full_noise)
        eng.search()
        steps = 0
        while any(orc.sims_done(g) < S + 1 for g in range(trees)):
            feats, plen = orc.think_select(K, rot[steps])
            pol, lg, val = ev.eval_batch(feats.reshape(K * trees, -1))
            orc.think_apply(pol.reshape(K, trees, A), lg.reshape(K, trees, A), val.reshape(K, trees), noise)
            steps += 1
        assert eng.think_steps() == steps
        r = eng.get_roots()
        for g in range(trees):
            b = orc.root(g)
            k = b["num_children"]
            assert r["root_count"][g] == S + 1 and r["num_children"][g] == k, (move, g)
            assert np.array_equal(r["action"][g, :k], b["action"][:k]), (move, g)
            assert np.array_equal(r["count"][g, :k], b["count"][:k]), (move, g, r["count"][g, :k], b["count"][:k])
            assert np.array_equal(r["mean"][g, :k].view(np.uint32), b["mean"][:k].view(np.uint32)), (move, g)
            assert np.array_equal(r["policy"][g, :k].view(np.uint32), b["policy"][:k].view(np.uint32)), (move, g)
        acts = np.full(eng.B, -1, np.int32)
        for g in range(trees):
            acts[g] = r["action"][g, int(r["count"][g].argmax())]
        res = eng.play_all(acts)
        for g in range(trees):
            assert res["applied"][g] == 1 and orc.play(g, int(acts[g])) == 1
            if res["terminal"][g]:
                eng.reset_game(g)
                orc.reset_game(g)
    eng.close()
    ev.close()


def test_think_mode_refuses_what_is_not_built():
    import minizero_b200
    with pytest.raises(minizero_b200.EngineError):
        engine(6, 6, 1, 16, think_batch_size=4, muzero=1, value_rescale=1)
    with pytest.raises(minizero_b200.EngineError):
        engine(1, 9, 1, 16, think_batch_size=4, use_gumbel=1)


@pytest.mark.parametrize("net,game,n,trees,S,K,moves", [("othello_mz_1bx32", 2, 8, 3, 40, 6, 4), ("go5_mz_1bx16", 1, 5, 2, 25, 4, 5)])
def test_on_device_muzero_think_search_matches_oracle(net, game, n, trees, S, K, moves):
    """MuZero think(): the root's initial inference as a batch of one lane, then K selections per step whose (parent hidden state, action) pairs go
    through the dynamics network together; hidden states live in their TREE's slots in evaluation order. Oracle fed by a network-only engine."""
    path = os.path.join(NETS, net + ".pt")
    if not os.path.exists(path):
        pytest.skip("net fixture missing")
    lib = oracle_lib.load()
    eng = engine(game, n, trees, S, think_batch_size=K, muzero=1)
    eng.load_network(path)
    ev = engine(game, n, trees * K, 2, muzero=1)
    ev.load_network(path)
    orc = oracle_lib.OracleSearch(lib, game, n, trees, S, muzero=1)
    rng = np.random.default_rng(11)
    A = eng.A
    for move in range(moves):
        noise = rng.dirichlet([0.3] * A, size=trees).astype(np.float32)
        full_noise = np.zeros((eng.B, A), np.float32)
        full_noise[:trees] = noise
        eng.set_search_inputs(None, full_noise)
        eng.search()
        store = [dict() for _ in range(trees)]  # tree -> evaluation slot -> hidden state
        steps = 0
        while any(orc.sims_done(g) < S + 1 for g in range(trees)):
            before = [orc.sims_done(g) for g in range(trees)]
            feats, plen = orc.think_select(K, None)
            pol, lg, val = np.zeros((K, trees, A), np.float32), np.zeros((K, trees, A), np.float32), np.zeros((K, trees), np.float32)
            lanes = [(k, g) for g in range(trees) for k in range(K) if plen[k, g] > 0]
            if steps == 0:
                assert all(k == 0 for k, _ in lanes) and len(lanes) == trees
                p, l, v, h = ev.eval_initial(np.stack([feats[0, g] for _, g in lanes]))
            else:
                leaves = [orc.think_leaf(k, g) for k, g in lanes]
                hidden = np.stack([store[g][ps] for (k, g), (ps, _) in zip(lanes, leaves)])
                p, l, v, h = ev.eval_recurrent(hidden, np.array([a for _, a in leaves], np.int32))
            queued = [0] * trees
            for j, (k, g) in enumerate(lanes):  # lanes are listed per tree in selection order: slots follow the finished simulations
                pol[k, g], lg[k, g], val[k, g] = p[j], l[j], v[j]
                store[g][before[g] + queued[g]] = h[j]
                queued[g] += 1
            orc.think_apply(pol, lg, val, noise)
            steps += 1
        assert eng.think_steps() == steps
        r = eng.get_roots()
        for g in range(trees):
            b = orc.root(g)
            k = b["num_children"]
            assert r["root_count"][g] == S + 1 and r["num_children"][g] == k, (move, g)
            assert np.array_equal(r["action"][g, :k], b["action"][:k]), (move, g)
            assert np.array_equal(r["count"][g, :k], b["count"][:k]), (move, g, r["count"][g, :k], b["count"][:k])
            assert np.array_equal(r["mean"][g, :k].view(np.uint32), b["mean"][:k].view(np.uint32)), (move, g)
        acts = np.full(eng.B, -1, np.int32)
        for g in range(trees):
            acts[g] = r["action"][g, int(r["count"][g].argmax())]
        res = eng.play_all(acts)
        for g in range(trees):
            assert res["applied"][g] == 1 and orc.play(g, int(acts[g])) == 1
            if res["terminal"][g]:
                eng.reset_game(g)
                orc.reset_game(g)
    eng.close()
    ev.close()
