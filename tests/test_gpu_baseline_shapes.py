"""Parity at the shapes BASELINE.json quotes ("what's weak" #1-2): the 20b x 256 19x19 network against TorchScript
fp32, whole searches through the fused tower at 19x19, configs[1] and configs[2] at full batch against the oracle on a sample
of the games, and a network with trained-scale statistics. Needs a B200: -m gpu."""
import os

import numpy as np
import pytest

import oracle_lib

pytestmark = pytest.mark.gpu

ROOT = oracle_lib.ROOT
NETS = os.path.join(ROOT, "oracle", "_ref", "nets")


def engine(*args, **kw):
    import minizero_b200
    return minizero_b200.Engine(*args, **kw)


def torchscript(net):
    torch = pytest.importorskip("torch")
    path = os.path.join(NETS, net + ".pt")
    if not os.path.exists(path):
        pytest.skip("net fixture missing (oracle/gen_nets.py needs the reference checkout)")
    return torch, torch.jit.load(path, map_location="cpu").eval(), path


def test_network_20bx256_19x19_matches_torchscript_fp32():
    """BASELINE configs[3] network (41 chained 3x3 conv layers, fp16 activations between them) at batch 128: logits, value and
    policy within the 1e-3 of north_star of the reference TorchScript module run in fp32 on the CPU"""
    torch, m, path = torchscript("go19_az_20bx256")
    batch, n = 128, 19
    eng = engine(1, n, batch, 4)
    eng.load_network(path)
    assert eng.conv_layers_per_launch() == 41  # one launch for the whole tower
    rng = np.random.default_rng(23)
    feats = (rng.random((batch, 18, n, n)) < 0.25).astype(np.float32)
    feats[0] = 0.0
    feats[1] = 1.0
    feats[2, :16] = 0.0  # the empty board with black to move: what every game starts from
    feats[2, 16], feats[2, 17] = 1.0, 0.0
    with torch.no_grad():
        ref = m(torch.from_numpy(feats))
    pol, lg, val = eng.eval_batch(feats)
    err = (np.abs(lg - ref["policy_logit"].numpy()).max(), np.abs(val - ref["value"].numpy().reshape(-1)).max(), np.abs(pol - ref["policy"].numpy()).max())
    print("NET-20BX256 19x19: max |d logit| %.2e, |d value| %.2e, |d policy| %.2e (logit range %.2f)" % (err + (float(np.abs(lg).max()),)))
    assert max(err) < 1e-3, err
    eng.close()


def torch_reference_forward(torch, dims, state, feats):
    """AlphaZeroNetwork.forward (network/py/alphazero_network.py:90-113, network_unit.py:6-65) written with torch.nn.functional in
    fp32 from a reference-named state_dict: the checker for weights the reference's create_network() did not draw"""
    F = torch.nn.functional
    t = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in state.items()}

    def cbn(x, conv, bn, pad):
        x = F.conv2d(x, t[conv + ".weight"], t[conv + ".bias"], padding=pad)
        return F.batch_norm(x, t[bn + ".running_mean"], t[bn + ".running_var"], t[bn + ".weight"], t[bn + ".bias"], training=False, eps=1e-5)

    with torch.no_grad():
        x = F.relu(cbn(torch.from_numpy(feats), "conv", "bn", 1))
        for b in range(dims["num_blocks"]):
            y = F.relu(cbn(x, f"residual_blocks.{b}.conv1", f"residual_blocks.{b}.bn1", 1))
            x = F.relu(cbn(y, f"residual_blocks.{b}.conv2", f"residual_blocks.{b}.bn2", 1) + x)
        p = F.relu(cbn(x, "policy.conv", "policy.bn", 0)).flatten(1)
        lg = F.linear(p, t["policy.fc.weight"], t["policy.fc.bias"])
        v = F.relu(cbn(x, "value.conv", "value.bn", 0)).flatten(1)
        v = F.relu(F.linear(v, t["value.fc1.weight"], t["value.fc1.bias"]))
        v = torch.tanh(F.linear(v, t["value.fc2.weight"], t["value.fc2.bias"]))
        return torch.softmax(lg, 1).numpy(), lg.numpy(), v.numpy().reshape(-1), float(x.abs().max())


def trained_scale_state(torch, dims, rng, gain, feats):
    """Weights with the statistics of a TRAINED net instead of an initialisation: every conv's output channels are scaled by
    10^U(-1, 1) (pre-BN variances spread over four decades, 1e-2 .. 1e2), and every BatchNorm's running mean / variance are
    then calibrated to the actual statistics of its input on `feats` (what training's moving averages converge to), layer by
    layer in fp32. bn2.weight = gain makes every block add a term of that size to the residual stream."""
    import __graft_entry__ as ge
    F = torch.nn.functional
    st = ge.make_random_state(dims, rng)

    def calibrate(x, conv, bn, pad, gamma=None):
        co = st[conv + ".weight"].shape[0]
        sc = (10.0 ** rng.uniform(-1, 1, size=co)).astype(np.float32)
        st[conv + ".weight"] = st[conv + ".weight"] * sc[:, None, None, None]
        st[conv + ".bias"] = st[conv + ".bias"] * sc
        y = F.conv2d(x, torch.from_numpy(st[conv + ".weight"]), torch.from_numpy(st[conv + ".bias"]), padding=pad)
        st[bn + ".running_mean"] = y.mean(dim=(0, 2, 3)).numpy().astype(np.float32)
        st[bn + ".running_var"] = y.var(dim=(0, 2, 3), unbiased=False).numpy().astype(np.float32)
        if gamma is not None:
            st[bn + ".weight"] = (gamma * (1 + 0.1 * rng.standard_normal(co))).astype(np.float32)
        return F.batch_norm(y, torch.from_numpy(st[bn + ".running_mean"]), torch.from_numpy(st[bn + ".running_var"]), torch.from_numpy(st[bn + ".weight"]),
                            torch.from_numpy(st[bn + ".bias"]), training=False, eps=1e-5)

    with torch.no_grad():
        x = F.relu(calibrate(torch.from_numpy(feats), "conv", "bn", 1))
        for b in range(dims["num_blocks"]):
            y = F.relu(calibrate(x, f"residual_blocks.{b}.conv1", f"residual_blocks.{b}.bn1", 1))
            x = F.relu(calibrate(y, f"residual_blocks.{b}.conv2", f"residual_blocks.{b}.bn2", 1, gamma=gain) + x)
        calibrate(x, "policy.conv", "policy.bn", 0)
        calibrate(x, "value.conv", "value.bn", 0)
    spread = [float(st[k].max() / max(st[k].min(), 1e-30)) for k in st if k.endswith("running_var") and st[k].size > 1]
    return st, max(spread)


@pytest.mark.parametrize("gain", [1.0, 10.0, 100.0])
def test_network_trained_scale_statistics(gain):
    """fp16 activations under trained-scale statistics (BatchNorm variances spread over four decades, every layer's output of unit
    variance, a residual stream growing to the tens / hundreds / thousands). Measured on B200 (round 2): logits of magnitude 3.1 -
    3.3 differ from the fp32 reference by 4.9e-3 / 6.6e-3 / 7.9e-3 at most (1.6 - 2.4e-3 of the largest logit), the value by
    1.6 - 2.1e-3, the policy PROBABILITIES by 2 - 4e-4. That is the limit of fp16 storage (11-bit significands on activations
    and weights, ~2e-4 relative per layer, 13 layers), not of this implementation: the absolute 1e-3 of north_star on the
    LOGITS holds for the random-init networks BASELINE.json names (logits of order 0.3: 1.9e-4 at 6b x 256, see
    test_network_20bx256_19x19_matches_torchscript_fp32 for 41 layers) and is asserted here on the policy probabilities; logits
    and value are held to 4e-3 of max(1, largest |logit|). DESIGN.md "Precision" discusses the split-precision alternative."""
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(31)
    n, batch, blocks = 9, 64, 6
    dims = dict(num_input_channels=18, input_height=n, input_width=n, num_hidden_channels=256, num_blocks=blocks, action_size=82, num_value_hidden_channels=256,
                discrete_value_size=1)
    feats = (rng.random((batch, 18, n, n)) < 0.3).astype(np.float32)
    st, spread = trained_scale_state(torch, dims, rng, gain, feats)
    pol_r, lg_r, val_r, act_max = torch_reference_forward(torch, dims, st, feats)
    eng = engine(1, n, batch, 4)
    eng.load_network((dims, st))
    pol, lg, val = eng.eval_batch(feats)
    scale = max(1.0, float(np.abs(lg_r).max()))
    err = (np.abs(lg - lg_r).max(), np.abs(val - val_r).max(), np.abs(pol - pol_r).max())
    print("TRAINED-SCALE gain %g: BN variance spread %.3g, largest activation %.3g, largest |logit| %.3g, max |d logit| %.2e, |d value| %.2e, |d policy| %.2e"
          % ((gain, spread, act_max, scale) + err))
    assert np.all(np.isfinite(lg)) and np.all(np.isfinite(val))
    assert err[0] < 4e-3 * scale and err[1] < 4e-3 and err[2] < 1e-3, err
    eng.close()


def test_network_saturates_instead_of_overflowing():
    """activations beyond the fp16 range: the reference (fp32) stays finite, so must this path (saturation at 65504, not inf / nan)"""
    import __graft_entry__ as ge
    rng = np.random.default_rng(37)
    n, batch = 9, 16
    dims = dict(num_input_channels=18, input_height=n, input_width=n, num_hidden_channels=128, num_blocks=2, action_size=82, num_value_hidden_channels=64,
                discrete_value_size=1)
    st = ge.make_random_state(dims, rng)
    st["bn.weight"] = (st["bn.weight"] * 3e5).astype(np.float32)  # the stem's output leaves the fp16 range
    eng = engine(1, n, batch, 4)
    eng.load_network((dims, st))
    pol, lg, val = eng.eval_batch((rng.random((batch, 18, n, n)) < 0.3).astype(np.float32))
    assert np.all(np.isfinite(lg)) and np.all(np.isfinite(val)) and np.all(np.isfinite(pol))
    eng.close()


def sampled_search_vs_oracle(game, n, B, S, net_path, moves, seed, sample, dirichlet_alpha=0.03, expect_wide=None):
    """The engine searches ALL B games (full BASELINE batch, one CUDA graph per move); the oracle re-runs the first `sample` games
    with the same rotations and noise, fed by a second, network-only engine (a position's network outputs do not depend on what
    else is in the batch: every row tile of the implicit GEMM accumulates its own rows in the same order). Root tables of the
    sampled games must agree bit for bit; the remaining games are held to the size-independent invariants."""
    lib = oracle_lib.load()
    eng = engine(game, n, B, S)
    eng.load_network(net_path)
    ev = engine(game, n, sample, 2)
    ev.load_network(net_path)
    if expect_wide is not None:  # the full batch runs conv_tower_wide_kernel, the sampled games' evaluations conv_tower_kernel: two kernels, one result
        assert eng.tower_is_wide() == expect_wide and ev.tower_is_wide() == 0
    orc = oracle_lib.OracleSearch(lib, game, n, sample, S)
    rng = np.random.default_rng(seed)
    A = eng.A
    for move in range(moves):
        rot = rng.integers(0, 8, size=(S + 1, B)).astype(np.uint8)
        noise = rng.dirichlet([dirichlet_alpha] * A, size=B).astype(np.float32)
        eng.set_search_inputs(rot, noise)
        eng.search()
        for c in range(S + 1):
            feats = orc.select(rot[c, :sample])
            pol, lg, val = ev.eval_batch(feats)
            orc.apply(pol, lg, val, noise[:sample])
        r = eng.get_roots()
        assert np.all(r["root_count"] == S + 1) and np.all(r["count"].sum(axis=1) == S) and np.all(np.isfinite(r["mean"]))
        for g in range(sample):
            b = orc.root(g)
            k = b["num_children"]
            assert r["num_children"][g] == k, (move, g)
            assert np.array_equal(r["action"][g, :k], b["action"][:k]), (move, g)
            assert np.array_equal(r["count"][g, :k], b["count"][:k]), (move, g, r["count"][g, :k], b["count"][:k])
            assert np.array_equal(r["mean"][g, :k].view(np.uint32), b["mean"][:k].view(np.uint32)), (move, g)
            assert np.array_equal(r["policy"][g, :k].view(np.uint32), b["policy"][:k].view(np.uint32)), (move, g)
        # every game plays its most visited move (first maximum, mcts.cpp:84-95); both sides stay in step
        actions = r["action"][np.arange(B), r["count"].argmax(axis=1)].astype(np.int32)
        res = eng.play_all(actions)
        assert np.all(res["applied"] == 1)
        for g in range(sample):
            assert orc.play(g, int(actions[g])) == 1
            assert bool(res["terminal"][g]) == orc.root_terminal(g)
        for g in np.nonzero(res["terminal"])[0]:
            eng.reset_game(int(g))
            if g < sample:
                orc.reset_game(int(g))
    eng.close()
    ev.close()


def test_config2_full_batch_sampled_games_match_oracle():
    """BASELINE configs[1]: 256 games x 400 simulations x 6b x 256, Dirichlet(0.03) + random rotations; 16 sampled games bit-exact"""
    torch, m, path = torchscript("go9_az_6bx256")
    sampled_search_vs_oracle(1, 9, 256, 400, path, moves=2, seed=41, sample=16, expect_wide=1)


def test_19x19_search_through_fused_tower_matches_oracle():
    """whole searches at 19x19 through the fused tower (176-row resident block, 4 weight stages): 4 games x 64 simulations x 5 moves"""
    torch, m, path = torchscript("go19_az_2bx128")
    sampled_search_vs_oracle(1, 19, 4, 64, path, moves=5, seed=43, sample=4, dirichlet_alpha=0.3)


def test_config4_net_search_sampled_games_match_oracle():
    """BASELINE configs[3] network (20b x 256) under a whole on-device search at 19x19: 8 games x 48 simulations, all games bit-exact"""
    torch, m, path = torchscript("go19_az_20bx256")
    sampled_search_vs_oracle(1, 19, 8, 48, path, moves=2, seed=47, sample=8)


def test_config3_full_batch_sampled_games_match_oracle():
    """BASELINE configs[2]: 512 games, Gumbel MuZero n=16 m=16, 3b x 128; 16 sampled games bit-exact incl. the Gumbel move decision"""
    torch, m, path = torchscript("othello_mz_3bx128")
    lib = oracle_lib.load()
    B, S, sample, game, n = 512, 16, 16, 2, 8
    opts = dict(use_gumbel=1, gumbel_noise=1, gumbel_sample_size=16)
    eng = engine(game, n, B, S, muzero=1, **opts)
    eng.load_network(path)
    ev = engine(game, n, sample, 2, muzero=1)
    ev.load_network(path)
    orc = oracle_lib.OracleSearch(lib, game, n, sample, S, muzero=1, **opts)
    rng = np.random.default_rng(53)
    A = eng.A
    for move in range(12):
        noise = rng.gumbel(size=(B, A)).astype(np.float32)
        eng.set_search_inputs(None, noise)
        eng.search()
        store = None
        for c in range(S + 1):
            feats = orc.select(None)
            if c == 0:
                pol, lg, val, hid = ev.eval_initial(feats)
                store = np.zeros((sample, S + 1) + hid.shape[1:], np.float32)
            else:
                parent = np.array([orc.leaf_parent_slot(g) for g in range(sample)])
                acts = np.array([orc.leaf_action(g) for g in range(sample)], np.int32)
                pol, lg, val, hid = ev.eval_recurrent(store[np.arange(sample), parent], acts)
            store[:, c] = hid
            orc.apply(pol, lg, val, noise[:sample])
        best = eng.gumbel_best_actions()
        r = eng.get_roots()
        assert np.all(r["root_count"] == S + 1) and np.all(r["count"].sum(axis=1) == S)
        for g in range(sample):
            b = orc.root(g)
            k = b["num_children"]
            assert r["num_children"][g] == k, (move, g)
            assert np.array_equal(r["action"][g, :k], b["action"][:k]), (move, g)
            assert np.array_equal(r["count"][g, :k], b["count"][:k]), (move, g)
            assert np.array_equal(r["mean"][g, :k].view(np.uint32), b["mean"][:k].view(np.uint32)), (move, g)
            assert np.array_equal(r["logit"][g, :k].view(np.uint32), b["logit"][:k].view(np.uint32)), (move, g)
            assert best[g] == orc.gumbel_best_action(g), (move, g)
        res = eng.play_all(best.astype(np.int32))
        assert np.all(res["applied"] == 1)
        for g in range(sample):
            assert orc.play(g, int(best[g])) == 1
            assert bool(res["terminal"][g]) == orc.root_terminal(g)
        for g in np.nonzero(res["terminal"])[0]:
            eng.reset_game(int(g))
            if g < sample:
                orc.reset_game(int(g))
    eng.close()
    ev.close()


def test_cooperative_tower_launch_gives_the_same_search():
    """mz_set_tower_cooperative: the tower launched with the cooperative attribute (co-residency guaranteed by the driver) inside the captured
    search graph; the root tables must be bit-identical to the plain cluster launch's"""
    torch, m, path = torchscript("go9_az_6bx256")
    rng = np.random.default_rng(5)
    B, S = 256, 40
    rot = rng.integers(0, 8, size=(S + 1, B)).astype(np.uint8)
    noise = rng.dirichlet([0.03] * 82, size=B).astype(np.float32)
    tables = []
    for coop in (0, 1):
        eng = engine(1, 9, B, S)
        eng.load_network(path)
        assert eng.tower_is_cooperative() == 0
        if coop:
            eng.set_tower_cooperative(True)
            assert eng.tower_is_cooperative() == 1
        eng.set_search_inputs(rot, noise)
        eng.search()
        r = eng.get_roots()
        tables.append((r["count"].copy(), r["mean"].copy(), r["policy"].copy()))
        eng.close()
    for a, b in zip(*tables):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_two_engines_on_one_device_with_cooperative_towers():
    """Two engines whose towers each need every SM, searching concurrently on ONE device from two host threads: with the cooperative launch the
    driver never lets the two grids share the SMs half and half (the case an ordinary cluster launch of a spin-waiting kernel can deadlock in — it
    would trap after 10 s). Runs in a subprocess under a timeout; both engines must finish and agree with a single-engine run bit for bit."""
    import subprocess
    import sys
    torch, m, path = torchscript("go9_az_6bx256")
    code = f'''
import sys, threading
import numpy as np
sys.path.insert(0, {oracle_lib.ROOT!r})
import minizero_b200 as mz
B, S = 256, 30
rng = np.random.default_rng(9)
rot = rng.integers(0, 8, size=(S + 1, B)).astype(np.uint8)
noise = rng.dirichlet([0.03] * 82, size=B).astype(np.float32)
def make(coop):
    e = mz.Engine(mz.GAME_GO, 9, B, S)
    e.load_network({path!r})
    if coop:
        e.set_tower_cooperative(True)
    return e
ref = make(False)
ref.set_search_inputs(rot, noise)
ref.search()
want = ref.get_roots()["count"].copy()
ref.close()
engines = [make(True), make(True)]
out = [None, None]
def run(i):
    for _ in range(6):
        engines[i].reset_game(-1)
        engines[i].set_search_inputs(rot, noise)
        engines[i].search()
    out[i] = engines[i].get_roots()["count"].copy()
threads = [threading.Thread(target=run, args=(i,)) for i in range(2)]
[t.start() for t in threads]
[t.join() for t in threads]
assert np.array_equal(out[0], want) and np.array_equal(out[1], want)
print("TWO_ENGINES_OK")
'''
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "TWO_ENGINES_OK" in r.stdout, r.stdout[-300:] + r.stderr[-800:]


def test_wide_and_narrow_towers_agree_bit_for_bit():
    """conv_tower_wide_kernel (two row tiles per CTA, chosen when a layer holds at least one wide unit per CTA pair: here 19x19 x 192 boards x 128 channels = 150
    units) against conv_tower_kernel (the same network in an engine of 48 boards): both accumulate K-block outer / tap inner, so every logit and value
    of a position must be identical whichever kernel — whichever batch — evaluated it"""
    import __graft_entry__ as ge
    rng = np.random.default_rng(17)
    dims = dict(num_input_channels=18, input_height=19, input_width=19, num_hidden_channels=128, num_blocks=3, action_size=362, num_value_hidden_channels=64,
                discrete_value_size=1)
    st = ge.make_random_state(dims, rng)
    feats = (rng.random((192, 18 * 361)) < 0.25).astype(np.float32)
    wide = engine(1, 19, 192, 2)
    wide.load_network((dims, st))
    assert wide.tower_is_wide() == 1
    pol_w, lg_w, val_w = wide.eval_batch(feats)
    wide.close()
    narrow = engine(1, 19, 48, 2)
    narrow.load_network((dims, st))
    assert narrow.tower_is_wide() == 0
    for lo in range(0, 192, 48):
        pol_n, lg_n, val_n = narrow.eval_batch(feats[lo:lo + 48])
        assert np.array_equal(lg_n.view(np.uint32), lg_w[lo:lo + 48].view(np.uint32)), lo
        assert np.array_equal(val_n.view(np.uint32), val_w[lo:lo + 48].view(np.uint32)), lo
        assert np.array_equal(pol_n.view(np.uint32), pol_w[lo:lo + 48].view(np.uint32)), lo
    narrow.close()
