"""Replays a golden case (tests/golden/*.npz, recorded from the compiled reference) through any
search engine exposing select/apply/root/play — the CPU oracle or the CUDA path — and checks, bit
for bit, the feature planes of every leaf evaluation and the root child table of every move."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def assert_tie_free(case):
    """std::sort is unstable; child order is only implementation-independent without exact policy ties."""
    gumbel = "actor_use_gumbel_noise=true" in str(case.get("conf", ""))
    for m in range(case["move_game"].size):
        k = int(case["move_num_children"][m])
        noise = case["child_noise"][m, :k]
        if gumbel:  # Gumbel noise goes to the logits; the priors stay sorted
            p = case["child_policy"][m, :k]
            assert np.all(p[:-1] > p[1:]), "exact policy tie in golden vector"
            continue
        if np.any(noise != 0):
            continue  # policies were mixed with noise after sorting; order not checkable from the table
        p = case["child_policy"][m, :k]
        assert np.all(p[:-1] > p[1:]), "exact policy tie in golden vector"


def replay(engine, case, check_features=True, on_move=None):
    A, F, S, B = (int(case[k]) for k in "AFSB")
    n_evals = case["eval_game"].size
    n_cycles = n_evals // B
    next_move = 0
    moves_of_game = {g: [m for m in range(case["move_game"].size) if case["move_game"][m] == g] for g in range(B)}
    move_ptr = {g: 0 for g in range(B)}
    checked_moves = 0
    for c in range(n_cycles):
        sl = slice(c * B, (c + 1) * B)
        assert np.all(case["eval_game"][sl] == np.arange(B))
        feats = engine.select(case["eval_rotation"][sl])
        muzero = "eval_leaf_action" in case
        if check_features:
            want = np.unpackbits(case["eval_features"][sl], axis=1)[:, :F].astype(np.float32)
            if muzero:  # only the initial inference consumes feature planes (zero_actor.cpp:59-61)
                root = case["eval_path_len"][sl] == 1
                assert np.array_equal(feats[root], want[root]), f"feature mismatch at cycle {c}"
            else:
                assert np.array_equal(feats, want), f"feature mismatch at cycle {c}"
        for g in range(B):
            assert engine.path_len(g) == case["eval_path_len"][sl][g], f"path length mismatch cycle {c} game {g}"
            if muzero:
                assert engine.leaf_action(g) == case["eval_leaf_action"][sl][g], f"leaf action mismatch cycle {c} game {g}"
                assert engine.path_hash(g) == case["eval_path_hash"][sl][g], f"path mismatch cycle {c} game {g}"
        use_noise = bool(np.any(case["child_noise"] != 0))  # actor_use_dirichlet_noise in the recording
        noise = np.zeros((B, A), np.float32) if use_noise else None
        for g in range(B if use_noise else 0):
            if engine.sims_done(g) == 0 and move_ptr[g] < len(moves_of_game[g]):
                noise[g] = case["child_noise"][moves_of_game[g][move_ptr[g]]]
        engine.apply(case["eval_policy"][sl], case["eval_logits"][sl], case["eval_value"][sl], noise)
        for g in range(B):
            if engine.sims_done(g) != S + 1:
                continue
            if move_ptr[g] >= len(moves_of_game[g]):
                return checked_moves  # recording stopped here
            m = moves_of_game[g][move_ptr[g]]
            move_ptr[g] += 1
            r = engine.root(g)
            k = int(case["move_num_children"][m])
            assert r["num_children"] == k
            assert r["root_count"] == case["root_count"][m] and r["root_mean"] == case["root_mean"][m] and r["root_value"] == case["root_value"][m]
            assert np.array_equal(r["action"][:k], case["child_action"][m, :k])
            for name in ("count", "mean", "policy", "logit", "noise", "value"):
                got, want = r[name][:k], case["child_" + name][m, :k]
                assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"root child {name} mismatch at move {m}"
            checked_moves += 1
            if on_move is not None:
                on_move(g, m, engine)
            if case["move_resign"][m]:
                engine.reset_game(g)
            else:
                assert engine.play(g, int(case["move_action"][m])) == 1
                if engine.root_terminal(g):
                    engine.reset_game(g)
    return checked_moves


def atari_planes(codes):
    """integer codes of a root record (RGB bytes, action ids; oracle/drivers/ref_stepper.cpp) -> the float planes of
    AtariEnv::getFeatures (atari.cpp:82,152-156): byte / 255.0f and id * 1.0f / 18, both single-precision divisions"""
    c = codes.reshape(32, 96 * 96).astype(np.float32)
    out = c / np.float32(255.0)
    out[0::4] = c[0::4] * np.float32(1.0) / np.float32(18)
    return out.reshape(-1)


def replay_atari(engine, case, check_features=True):
    """Atari MuZero recordings: the emulator's frames come from the recorded root planes (the newest history entry of the next root
    evaluation of the game), rewards from the recorded network outputs; checks planes, paths, root tables incl. the children's
    rewards and the value bounds the search ended with"""
    A, S, B = (int(case[k]) for k in "ASB")
    n_cycles = case["eval_game"].size // B
    roots_of_game = {g: [r for r in range(case["root_game"].size) if case["root_game"][r] == g] for g in range(B)}
    root_ptr = {g: 0 for g in range(B)}
    moves_of_game = {g: [m for m in range(case["move_game"].size) if case["move_game"][m] == g] for g in range(B)}
    move_ptr = {g: 0 for g in range(B)}
    use_noise = bool(np.any(case["child_noise"] != 0))

    def newest(r):  # (action id of the newest history entry, its frame)
        planes = case["root_planes"][r].reshape(32, 96 * 96)
        return int(planes[28, 0]), planes[29:32]

    for g in range(B):
        engine.observe(g, -1, newest(roots_of_game[g][0])[1])
    checked = 0
    for c in range(n_cycles):
        sl = slice(c * B, (c + 1) * B)
        feats = engine.select(None)
        for g in range(B):
            assert engine.path_len(g) == case["eval_path_len"][sl][g], f"path length mismatch cycle {c} game {g}"
            assert engine.leaf_action(g) == case["eval_leaf_action"][sl][g], f"leaf action mismatch cycle {c} game {g}"
            assert engine.path_hash(g) == case["eval_path_hash"][sl][g], f"path mismatch cycle {c} game {g}"
            if case["eval_path_len"][sl][g] == 1:
                r = roots_of_game[g][root_ptr[g]]
                assert case["root_cycle"][r] == c
                if check_features and feats is not None:
                    want = atari_planes(case["root_planes"][r])
                    assert np.array_equal(feats[g].view(np.uint32), want.view(np.uint32)), f"plane mismatch at cycle {c} game {g}"
        noise = np.zeros((B, A), np.float32) if use_noise else None
        for g in range(B if use_noise else 0):
            if engine.sims_done(g) == 0 and move_ptr[g] < len(moves_of_game[g]):
                noise[g] = case["child_noise"][moves_of_game[g][move_ptr[g]]]
        engine.apply(case["eval_policy"][sl], case["eval_logits"][sl], case["eval_value"][sl], noise, reward=case["eval_reward"][sl])
        for g in range(B):
            if engine.sims_done(g) != S + 1:
                continue
            if move_ptr[g] >= len(moves_of_game[g]):
                return checked
            m = moves_of_game[g][move_ptr[g]]
            move_ptr[g] += 1
            r = engine.root(g)
            k = int(case["move_num_children"][m])
            assert r["num_children"] == k
            assert r["root_count"] == case["root_count"][m] and r["root_mean"] == case["root_mean"][m] and r["root_value"] == case["root_value"][m]
            assert np.array_equal(r["action"][:k], case["child_action"][m, :k])
            for name in ("count", "mean", "policy", "logit", "noise", "value", "reward"):
                got, want = r[name][:k], case["child_" + name][m, :k]
                assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"root child {name} mismatch at move {m}: {got} {want}"
            assert r["bound_size"] == case["bound_size"][m], (m, r["bound_size"], case["bound_size"][m])
            assert np.float32(r["bound_lo"]) == case["bound_lo"][m] and np.float32(r["bound_hi"]) == case["bound_hi"][m], f"value bounds mismatch at move {m}"
            checked += 1
            root_ptr[g] += 1
            if root_ptr[g] >= len(roots_of_game[g]):
                return checked  # the recording holds no later frame of this game
            action, frame = newest(roots_of_game[g][root_ptr[g]])
            if case["move_resign"][m] or case["env_terminal"][m]:
                engine.reset_game(g)
                engine.observe(g, -1, frame)
            else:
                assert engine.play(g, int(case["move_action"][m])) == 1
                assert action == case["move_action"][m]
                engine.observe(g, action, frame)
    return checked


def replay_think(engine, case, features_of_duplicates=True):
    """Console-search recordings (oracle/gen_think_golden.py; one tree, actor_mcts_think_batch_size = K): every batched step's selections must agree
    (which lanes exist, path lengths, which leaves are duplicates, the planes), the recorded network outputs are applied, and every search must end
    with the recorded root child table. `engine` exposes think_select / think_apply / root / play / sims_done with lane-major arrays [K][1]."""
    A, F, S, K = (int(case[k]) for k in "AFSK")
    searches = int(case["move_action"].size)
    checked = 0
    for m in range(searches):
        steps = np.unique(case["sel_step"][case["sel_search"] == m])
        use_noise = bool(np.any(case["child_noise"][m] != 0))
        for st in steps:
            idx = np.nonzero(case["sel_step"] == st)[0]
            n = idx.size
            assert np.array_equal(case["sel_lane"][idx], np.arange(n))
            rot = np.zeros((K, 1), np.uint8)
            rot[:n, 0] = case["sel_rotation"][idx]
            first = (engine.sims_done(0) == 0)
            feats, plen = engine.think_select(K, rot)
            want_len = np.zeros(K, np.int64)
            want_len[:n] = np.where(case["sel_leaf_vloss"][idx] == 0, case["sel_path_len"][idx], -case["sel_path_len"][idx])
            assert np.array_equal(plen[:, 0], want_len), (m, st, plen[:, 0], want_len)
            want = np.unpackbits(case["sel_features"][idx], axis=1)[:, :F].astype(np.float32)
            muzero = "nn_type_name=muzero" in str(case["conf"])
            for k in range(n):
                if muzero and abs(int(want_len[k])) != 1:
                    continue  # below the root a MuZero search pushes (hidden state, action), no planes (zero_actor.cpp:62-66)
                if want_len[k] > 0 or features_of_duplicates:
                    assert np.array_equal(feats[k, 0], want[k]), f"planes differ: search {m} step {st} lane {k}"
            pol, lg, val = np.zeros((K, 1, A), np.float32), np.zeros((K, 1, A), np.float32), np.zeros((K, 1), np.float32)
            oidx = np.nonzero(case["out_step"] == st)[0]
            assert np.array_equal(case["out_lane"][oidx], np.nonzero(want_len > 0)[0])
            for o in oidx:
                k = int(case["out_lane"][o])
                pol[k, 0], lg[k, 0], val[k, 0] = case["out_policy"][o], case["out_logits"][o], case["out_value"][o]
            noise = case["child_noise"][m][None, :].astype(np.float32) if (use_noise and first) else None
            engine.think_apply(pol, lg, val, noise)
        assert engine.sims_done(0) == S + 1
        r = engine.root(0)
        k = int(case["move_num_children"][m])
        assert r["num_children"] == k
        assert r["root_count"] == case["root_count"][m] and r["root_mean"] == case["root_mean"][m] and r["root_value"] == case["root_value"][m]
        assert np.array_equal(r["action"][:k], case["child_action"][m, :k])
        for name in ("count", "mean", "policy", "logit", "noise", "value"):
            got, want = r[name][:k], case["child_" + name][m, :k]
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"root child {name} differs after search {m}"
        assert np.all(case["child_vloss"][m] == 0)
        checked += 1
        assert engine.play(0, int(case["move_action"][m])) == 1
    return checked
