"""Edge cases of the C ABI on the device: argument and state errors (status code + message, never a crash), smallest sizes,
ragged batches, model reload, illegal moves. Needs a B200: -m gpu."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib

pytestmark = pytest.mark.gpu
NETS = os.path.join(oracle_lib.ROOT, "oracle", "_ref", "nets")


def mz():
    import minizero_b200
    return minizero_b200


def net(name):
    path = os.path.join(NETS, name + ".pt")
    if not os.path.exists(path):
        pytest.skip("net fixture missing")
    return path


@pytest.mark.parametrize("kw,needle", [
    (dict(game=8, board_size=9), "unsupported game"),
    (dict(game=7, board_size=9), "7 x 7"),
    (dict(game=1, board_size=1), "board_size"),
    (dict(game=1, board_size=20), "board_size"),
    (dict(game=2, board_size=7), "othello"),
    (dict(game=1, board_size=9, num_games=0), "positive"),
    (dict(game=1, board_size=9, num_simulation=0), "positive"),
    (dict(game=2, board_size=8, use_gumbel=1, gumbel_sample_size=1), "gumbel"),
    (dict(game=1, board_size=9, device=99), "device"),
    (dict(game=1, board_size=19, num_simulation=4000), "2^20"),
])
def test_create_refuses_bad_configurations(kw, needle):
    args = dict(game=1, board_size=9, num_games=2, num_simulation=4)
    args.update(kw)
    with pytest.raises(mz().EngineError) as ei:
        mz().Engine(args.pop("game"), args.pop("board_size"), args.pop("num_games"), args.pop("num_simulation"), **args)
    assert needle in str(ei.value)


def test_state_errors_are_reported():
    m = mz()
    eng = m.Engine(m.GAME_TICTACTOE, 3, 2, 4)
    with pytest.raises(m.EngineError, match="not finalized"):
        eng.search()
    with pytest.raises(m.EngineError, match="not finalized"):
        eng.eval_batch(np.zeros((1, 4, 3, 3), np.float32))
    eng.load_network(net("ttt_az_2bx32"))
    with pytest.raises(m.EngineError, match="muzero"):
        eng.eval_recurrent(np.zeros((1, 32 * 9), np.float32), np.zeros(1, np.int32))  # an AlphaZero engine has no dynamics network
    with pytest.raises(m.EngineError, match="gumbel"):
        eng.gumbel_best_actions()
    with pytest.raises(m.EngineError, match="num_evals"):
        eng.search(num_evals=6)
    # a network of another game / type is refused and the engine stays usable
    with pytest.raises(m.EngineError, match="do not match|does not match"):
        eng.load_network(net("go9_az_1bx16"))
    oth = m.Engine(m.GAME_OTHELLO, 8, 2, 4)  # an AlphaZero-type engine must refuse a muzero network of the right shape
    with pytest.raises(m.EngineError, match="does not match"):
        oth.load_network(net("othello_mz_1bx32"))
    oth.close()
    eng.load_network(net("ttt_az_2bx32"))
    eng.search()
    assert np.all(eng.get_roots()["root_count"] == 5)
    eng.close()
    mzeng = m.Engine(m.GAME_OTHELLO, 8, 2, 4, muzero=1)
    mzeng.load_network(net("othello_mz_1bx32"))
    with pytest.raises(m.EngineError, match="whole"):
        mzeng.search(num_evals=2)
    mzeng.close()


def test_smallest_engine_and_partial_searches():
    """one game, one simulation; and an AlphaZero search split into two partial runs equals the whole one"""
    m = mz()
    eng = m.Engine(m.GAME_TICTACTOE, 3, 1, 1)
    eng.load_network(net("ttt_az_2bx32"))
    eng.search()
    r = eng.get_roots()
    assert r["root_count"][0] == 2 and r["num_children"][0] == 9 and r["count"][0].sum() == 1
    eng.close()
    a = m.Engine(m.GAME_GO, 9, 4, 24)
    b = m.Engine(m.GAME_GO, 9, 4, 24)
    for e in (a, b):
        e.load_network(net("go9_az_1bx16"))
        e.set_search_inputs(None, None)
    a.search()
    b.search(num_evals=10)
    assert np.all(b.get_roots()["root_count"] == 10)
    b.search(num_evals=15)
    ra, rb = a.get_roots(), b.get_roots()
    assert np.array_equal(ra["count"], rb["count"]) and np.array_equal(ra["mean"].view(np.uint32), rb["mean"].view(np.uint32))
    a.close(), b.close()


def test_ragged_batches_and_reload_give_the_same_outputs():
    m = mz()
    eng = m.Engine(m.GAME_GO, 9, 16, 4)
    path = net("go9_az_2bx64")
    eng.load_network(path)
    rng = np.random.default_rng(0)
    feats = (rng.random((16, 18, 9, 9)) < 0.3).astype(np.float32)
    full = eng.eval_batch(feats)
    one = eng.eval_batch(feats[:1])
    five = eng.eval_batch(feats[:5])
    for k in range(3):
        assert np.array_equal(full[k][:1], one[k]) and np.array_equal(full[k][:5], five[k])  # a position's outputs do not depend on its batch
    with pytest.raises(m.EngineError):
        eng.eval_batch(np.zeros((17, 18, 9, 9), np.float32))  # more positions than games
    eng.load_network(path)  # load_model of the next iteration: same shape, buffers and graphs stay
    again = eng.eval_batch(feats)
    for k in range(3):
        assert np.array_equal(full[k], again[k])
    eng.close()


def test_illegal_moves_are_refused_and_leave_the_game_untouched():
    m = mz()
    eng = m.Engine(m.GAME_OTHELLO, 8, 2, 4, muzero=1)
    res = eng.play_all(np.array([0, 64], np.int32))  # a corner and a pass are both illegal in the opening position
    assert np.all(res["applied"] == 0)
    res = eng.play_all(np.array([20, -1], np.int32))  # -1: no move for that game
    assert res["applied"][0] == 1 and res["turn"][0] == 2 and res["num_legal"][0] == 3
    assert res["applied"][1] == 0 and res["turn"][1] == 1
    go = m.Engine(m.GAME_GO, 5, 1, 4)
    assert go.play(0, 12) == 1 and go.play(0, 12) == 0 and go.play(0, 25) == 1 and go.play(0, 99) == 0 and go.play(0, 25) == 1
    assert go.root_terminal(0)  # two passes in a row (go.cpp:249-251)
    eng.close(), go.close()
