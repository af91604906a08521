"""Small driver for compute-sanitizer (memcheck / racecheck / initcheck / synccheck): one short on-device search of every kind
(AlphaZero Go with captures and rotations + noise, Gumbel MuZero Othello, console think() lanes, the wide tower) plus the per-phase hooks, all through the C ABI."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import minizero_b200 as mz  # noqa: E402

NETS = os.path.join(ROOT, "oracle", "_ref", "nets")
rng = np.random.default_rng(0)

eng = mz.Engine(mz.GAME_GO, 5, 4, 24)
eng.load_network(os.path.join(NETS, "go5_az_1bx16.pt"))
for move in range(6):
    eng.set_search_inputs(rng.integers(0, 8, size=(25, 4)).astype(np.uint8), rng.dirichlet([0.3] * 26, size=4).astype(np.float32))
    eng.search()
    r = eng.get_roots()
    assert np.all(r["root_count"] == 25)
    eng.play_max_count(auto_reset=True, read_back=True)
eng.close()

eng = mz.Engine(mz.GAME_OTHELLO, 8, 4, 16, muzero=1, use_gumbel=1, gumbel_noise=1, gumbel_sample_size=16)
eng.load_network(os.path.join(NETS, "othello_mz_1bx32.pt"))
for move in range(6):
    eng.set_search_inputs(None, rng.gumbel(size=(4, 65)).astype(np.float32))
    eng.search()
    assert np.all(eng.get_roots()["root_count"] == 17)
    eng.play_max_count(auto_reset=True, read_back=True)
eng.close()

# console think(): K selections per tree and step under virtual loss (lane views of the state)
eng = mz.Engine(mz.GAME_GO, 5, 2, 30, think_batch_size=6)
eng.load_network(os.path.join(NETS, "go5_az_1bx16.pt"))
for move in range(3):
    eng.set_search_inputs(rng.integers(0, 8, size=(31, 12)).astype(np.uint8), np.zeros((12, 26), np.float32))
    eng.search()
    assert np.all(eng.get_roots()["root_count"][:2] == 31)
    acts = np.full(eng.B, -1, np.int32)
    r = eng.get_roots()
    for g in range(2):
        acts[g] = r["action"][g, int(r["count"][g].argmax())]
    eng.play_all(acts)
eng.close()

# the wide tower (two row tiles per CTA): 19x19 with enough boards for two wide units per CTA pair and layer, a 2-block x 128-channel net
if os.environ.get("SAN_WIDE", "1") == "1":
    import __graft_entry__ as ge  # noqa: E402
    dims = dict(num_input_channels=18, input_height=19, input_width=19, num_hidden_channels=128, num_blocks=2, action_size=362, num_value_hidden_channels=64,
                discrete_value_size=1)
    eng = mz.Engine(mz.GAME_GO, 19, 192, 2)
    eng.load_network((dims, ge.make_random_state(dims, rng)))
    feats = (rng.random((192, 18 * 361)) < 0.2).astype(np.float32)
    pol, lg, val = eng.eval_batch(feats)
    assert np.all(np.isfinite(lg)) and np.all(np.isfinite(val))
    eng.close()

# the narrow tower (conv_tower_kernel: TMA-store epilogue, residual prefetch behind one counter acquire): 9x9 boards, a 2-block x 128-channel net, a whole search
if os.environ.get("SAN_NARROW", "1") == "1":
    import __graft_entry__ as ge  # noqa: E402
    dims = dict(num_input_channels=18, input_height=9, input_width=9, num_hidden_channels=128, num_blocks=2, action_size=82, num_value_hidden_channels=64,
                discrete_value_size=1)
    eng = mz.Engine(mz.GAME_GO, 9, 8, 12)
    eng.load_network((dims, ge.make_random_state(dims, rng)))
    assert eng.tower_is_wide() == 0 and eng.conv_layers_per_launch() == 5
    eng.set_search_inputs(rng.integers(0, 8, size=(13, 8)).astype(np.uint8), rng.dirichlet([0.3] * 82, size=8).astype(np.float32))
    eng.search()
    assert np.all(eng.get_roots()["root_count"] == 13)
    eng.close()

import golden_replay  # noqa: E402
case = golden_replay.load_case("ttt_s50_b2")
eng = mz.Engine(mz.GAME_TICTACTOE, 3, 2, 50)
print("replayed", golden_replay.replay(eng, case), "moves")
eng.close()
print("sanitizer driver done")
