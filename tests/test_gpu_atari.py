"""Atari MuZero (BASELINE configs[4]) on the device: tree kernels against recordings of the -DATARI reference, the muzero_atari
network against TorchScript fp32, whole on-device searches against the oracle. Needs a B200: -m gpu."""
import os

import numpy as np
import pytest

import golden_replay
import oracle_lib

pytestmark = pytest.mark.gpu

NETS = os.path.join(oracle_lib.ROOT, "oracle", "_ref", "nets")
ATARI_SEARCH = dict(muzero=1, value_rescale=1, reward_discount=0.997)


def engine(*args, **kw):
    import minizero_b200
    return minizero_b200.Engine(*args, **kw)


def torchscript(net):
    torch = pytest.importorskip("torch")
    path = os.path.join(NETS, net + ".pt")
    if not os.path.exists(path):
        pytest.skip("net fixture missing (oracle/gen_nets.py needs the reference checkout)")
    return torch, torch.jit.load(path, map_location="cpu").eval(), path


@pytest.mark.parametrize("name", ["atari_mz_s20_b2", "atari_mz_s50_b2_det", "atari_mz_s18_gumbel_b2"])
def test_tree_kernels_replay_reference_recording_atari(name):
    """bit-exact: the 32 planes of every root, paths, root child tables incl. rewards, the value bounds every search ended with"""
    case = golden_replay.load_case(name)
    eng = engine(6, 6, int(case["B"]), int(case["S"]), **oracle_lib.conf_overrides(case["conf"]))
    checked = golden_replay.replay_atari(eng, case)
    assert checked >= case["move_game"].size - int(case["B"])
    eng.close()


def scalar_from_bins(probs):
    """MuZeroNetwork::forward (network/muzero_network.h:157-171): sum_i p_i * (i - 300) accumulated in bin order in f32, then
    utils::invertValue (utils/utils.h:102-108)"""
    p = probs.astype(np.float32)
    n = p.shape[1]
    acc = np.zeros(p.shape[0], np.float32)
    for i in range(n):
        acc = (acc + p[:, i] * np.float32(i - n // 2)).astype(np.float32)
    eps = np.float32(0.001)
    r = (np.sqrt(1 + 4 * eps * (np.abs(acc) + 1 + eps)) - 1) / (2 * eps)
    return (np.sign(acc) * (r * r - 1)).astype(np.float32)


def synthetic_planes(rng, batch):
    """planes as AtariEnv::getFeatures lays them out: per history entry an action plane (id / 18) and three colour planes (byte / 255)"""
    f = np.zeros((batch, 32, 96, 96), np.float32)
    for i in range(8):
        f[:, 4 * i] = (rng.integers(0, 18, size=batch).astype(np.float32) / np.float32(18))[:, None, None]
        f[:, 4 * i + 1:4 * i + 4] = rng.integers(0, 256, size=(batch, 3, 96, 96)).astype(np.float32) / np.float32(255)
    f[0, :28] = 0.0  # a game's first position: seven empty history entries
    return f


@pytest.mark.parametrize("net,batch", [("atari_mz_1bx32", 8), ("atari_mz_1bx256", 16)])
def test_atari_network_matches_torchscript_fp32(net, batch):
    """initial_inference and recurrent_inference of MuZeroAtariNetwork (network/py/muzero_atari_network.py:157-183): stride-2 convolutions,
    average pooling, 18 action planes, reward head on the unscaled dynamics output, 601-bin heads with expectation + invertValue.
    Policy logits within 1e-3; scaled hidden state (fp16 here) within 2e-3; value / reward within 1e-3 of their magnitude"""
    torch, m, path = torchscript(net)
    eng = engine(6, 6, batch, 4, **ATARI_SEARCH)
    eng.load_network(path)
    rng = np.random.default_rng(61)
    feats = synthetic_planes(rng, batch)
    with torch.no_grad():
        ref = m.initial_inference(torch.from_numpy(feats))
    pol, lg, val, hid = eng.eval_initial(feats)
    ref_val = scalar_from_bins(ref["value"].numpy())
    ref_hid = ref["hidden_state"].numpy().reshape(batch, -1)
    print("ATARI-NET %s initial: max |d logit| %.2e, |d policy| %.2e, |d hidden| %.2e, |d value| %.2e (|value| up to %.3g)"
          % (net, np.abs(lg - ref["policy_logit"].numpy()).max(), np.abs(pol - ref["policy"].numpy()).max(), np.abs(hid - ref_hid).max(), np.abs(val - ref_val).max(), np.abs(ref_val).max()))
    assert np.abs(lg - ref["policy_logit"].numpy()).max() < 1e-3
    assert np.abs(pol - ref["policy"].numpy()).max() < 1e-3
    assert hid.min() >= 0.0 and hid.max() <= 1.0 and np.abs(hid - ref_hid).max() < 2e-3
    assert np.all(np.abs(val - ref_val) < 1e-3 * np.maximum(1.0, np.abs(ref_val)))
    # recurrent inference from the REFERENCE's hidden states, every action id represented
    actions = (np.arange(batch) % 18).astype(np.int32)
    planes = np.zeros((batch, 18, 6, 6), np.float32)
    planes[np.arange(batch), actions] = 1.0
    with torch.no_grad():
        ref2 = m.recurrent_inference(ref["hidden_state"], torch.from_numpy(planes))
    pol2, lg2, val2, hid2 = eng.eval_recurrent(ref_hid, actions)
    rew2 = eng.eval_rewards(batch)
    ref_val2, ref_rew2 = scalar_from_bins(ref2["value"].numpy()), scalar_from_bins(ref2["reward"].numpy())
    print("ATARI-NET %s recurrent: max |d logit| %.2e, |d hidden| %.2e, |d value| %.2e, |d reward| %.2e (|reward| up to %.3g)"
          % (net, np.abs(lg2 - ref2["policy_logit"].numpy()).max(), np.abs(hid2 - ref2["hidden_state"].numpy().reshape(batch, -1)).max(), np.abs(val2 - ref_val2).max(),
             np.abs(rew2 - ref_rew2).max(), np.abs(ref_rew2).max()))
    assert np.abs(lg2 - ref2["policy_logit"].numpy()).max() < 1e-3
    assert np.abs(pol2 - ref2["policy"].numpy()).max() < 1e-3
    assert np.abs(hid2 - ref2["hidden_state"].numpy().reshape(batch, -1)).max() < 2e-3
    assert np.all(np.abs(val2 - ref_val2) < 1e-3 * np.maximum(1.0, np.abs(ref_val2)))
    assert np.all(np.abs(rew2 - ref_rew2) < 1e-3 * np.maximum(1.0, np.abs(ref_rew2)))
    eng.close()


def run_atari_search_vs_oracle(B, S, net_path, moves, seed, **opts):
    """whole-move on-device Atari MuZero searches (CUDA graph: tree step, screen ring -> planes, the four representation stages with
    their space-to-depth / pooling kernels, dynamics tower, reward / value / policy heads, hidden-state scaling) against the oracle
    fed with the engine's own network outputs, move after move; random screens stand in for the emulator on both sides"""
    lib = oracle_lib.load()
    eng = engine(6, 6, B, S, **ATARI_SEARCH, **opts)
    eng.load_network(net_path)
    ev = engine(6, 6, B, 2, **ATARI_SEARCH)  # network-only engine
    ev.load_network(net_path)
    orc = oracle_lib.OracleSearch(lib, oracle_lib.GAME_ATARI, 6, B, S, **ATARI_SEARCH, **opts)
    rng = np.random.default_rng(seed)
    A = eng.A
    gumbel = bool(opts.get("use_gumbel"))
    legal = [a for a in range(18) if (oracle_lib.ATARI_LEGAL_MASK >> a) & 1]

    def new_frames():
        f = np.zeros((B, 3, 96, 96), np.uint8)
        f[:, :, ::8, ::8] = rng.integers(0, 256, size=(B, 3, 12, 12))  # sparse: cheap to draw, every plane still differs
        f[:, 2] += rng.integers(0, 40, size=(B, 1, 1)).astype(np.uint8)
        return f

    frames = new_frames()
    eng.observe_all(np.full(B, -1, np.int32), frames)
    for g in range(B):
        orc.observe(g, -1, frames[g])
    for move in range(moves):
        noise = (rng.gumbel(size=(B, A)) if opts.get("gumbel_noise") else rng.dirichlet([0.3] * A, size=B)).astype(np.float32)
        eng.set_search_inputs(None, noise)
        eng.search()
        store = None
        for c in range(S + 1):
            feats = orc.select(None)
            if c == 0:
                pol, lg, val, hid = ev.eval_initial(feats)
                rew = np.zeros(B, np.float32)
                store = np.zeros((B, S + 1) + hid.shape[1:], np.float32)
            else:
                parent = np.array([orc.leaf_parent_slot(g) for g in range(B)])
                acts = np.array([orc.leaf_action(g) for g in range(B)], np.int32)
                pol, lg, val, hid = ev.eval_recurrent(store[np.arange(B), parent], acts)
                rew = ev.eval_rewards(B)
            store[:, c] = hid
            orc.apply(pol, lg, val, noise, reward=rew)
        best = eng.gumbel_best_actions() if gumbel else None
        actions = np.zeros(B, np.int32)
        for g in range(B):
            a, b = eng.root(g), orc.root(g)
            assert a["num_children"] == b["num_children"] == len(legal), (move, g)
            k = a["num_children"]
            assert np.array_equal(a["action"][:k], b["action"][:k]), (move, g)
            assert np.array_equal(a["count"][:k], b["count"][:k]), (move, g, a["count"][:k], b["count"][:k])
            for name in ("mean", "logit", "reward", "value"):
                assert np.array_equal(a[name][:k].view(np.uint32), b[name][:k].view(np.uint32)), (move, g, name)
            assert a["bound_size"] == b["bound_size"] and np.float32(a["bound_lo"]) == np.float32(b["bound_lo"]) and np.float32(a["bound_hi"]) == np.float32(b["bound_hi"]), (move, g)
            assert a["root_count"] == S + 1 and a["count"][:k].sum() == S
            if gumbel:
                assert best[g] == orc.gumbel_best_action(g), (move, g)
                actions[g] = best[g]
            else:
                actions[g] = a["action"][int(np.argmax(a["count"][:k]))]
        res = eng.play_all(actions)
        frames = new_frames()
        eng.observe_all(actions, frames)
        for g in range(B):
            assert res["applied"][g] == 1 and orc.play(g, int(actions[g])) == 1
            orc.observe(g, int(actions[g]), frames[g])
    eng.close()
    ev.close()


def test_on_device_atari_search_matches_oracle():
    torch, m, path = torchscript("atari_mz_1bx32")
    run_atari_search_vs_oracle(4, 50, path, moves=10, seed=71)


def test_on_device_atari_gumbel_search_matches_oracle():
    torch, m, path = torchscript("atari_mz_1bx32")
    run_atari_search_vs_oracle(4, 18, path, moves=10, seed=73, use_gumbel=1, gumbel_noise=1, gumbel_sample_size=8, gumbel_sigma_scale_c=0.1)


def test_on_device_atari_search_matches_oracle_1bx256():
    """the reference's default network size (1 block x 256 channels, configuration.cpp:70-71) at a BASELINE-like batch slice"""
    torch, m, path = torchscript("atari_mz_1bx256")
    run_atari_search_vs_oracle(8, 50, path, moves=3, seed=79)
