"""search_core.cuh (the source of the CUDA tree/env kernels) compiled for the host with a 1-lane warp,
replayed against the recordings of the compiled reference. Catches logic errors without a GPU; the
-m gpu tests repeat the same replay through the real kernels and the C-ABI."""
import pytest

import golden_replay
import hostsim_lib

import oracle_lib

CASES = {"ttt_s50_b2": (0, 3), "ttt_s50_b1_det": (0, 3), "go5_s24_b2": (1, 5), "go9_s32_b2": (1, 9), "go19_s8_b2": (1, 19), "nogo9_s8_b2": (3, 9), "gomoku15_s8_b2": (4, 15), "hex11_s8_b2": (5, 11), "killallgo7_s16_b2": (7, 7),
         "go5_mz_s16_b2": (1, 5), "ttt_gmz_s16_b2": (0, 3), "othello_gmz_s16_b2": (2, 8), "othello_gmz_s32_m8_b2": (2, 8), "othello_mz_s24_b2": (2, 8),
         # odd Gumbel sample sizes and BASELINE search lengths: 400 simulations with the 6b x 256 net, 19x19 with 800 simulations
         "go5_gmz_s64_m12_b2": (1, 5), "go5_gmz_s100_m14_b2": (1, 5), "go9_s400_b2": (1, 9), "go19_s800_b2": (1, 19)}


@pytest.mark.parametrize("name", list(CASES))
def test_search_core_matches_reference_recording(name):
    game, n = CASES[name]
    case = golden_replay.load_case(name)
    eng = hostsim_lib.HostSimSearch(hostsim_lib.load(), game, n, int(case["B"]), int(case["S"]), **oracle_lib.conf_overrides(case["conf"]))
    compared = []

    def on_move(g, m, engine):  # the finished tree of every move: all per-level selection routines must agree on every node
        r = engine.check_level_variants(g)
        assert r >= 0, f"selection variants disagree on node {-r - 1} (move {m})"
        compared.append(r)

    checked = golden_replay.replay(eng, case, on_move=on_move)
    assert checked >= case["move_game"].size - int(case["B"])
    assert sum(compared) > 0


@pytest.mark.parametrize("name", ["atari_mz_s20_b2", "atari_mz_s50_b2_det", "atari_mz_s18_gumbel_b2"])
def test_search_core_matches_reference_recording_atari(name):
    """Atari MuZero tree semantics of search_core.cuh (rewards, value-bound table, rescaled Q, ATARI init-Q) against the -DATARI reference"""
    case = golden_replay.load_case(name)
    eng = hostsim_lib.HostSimSearch(hostsim_lib.load(), 6, 6, int(case["B"]), int(case["S"]), **oracle_lib.conf_overrides(case["conf"]))
    checked = golden_replay.replay_atari(eng, case, check_features=False)
    assert checked >= case["move_game"].size - int(case["B"])
    for g in range(int(case["B"])):
        assert eng.check_level_variants(g) >= 0


def test_candidate_sort_is_libstdcxx_std_sort():
    """mz_std_sort_candidates (search_core.cuh) and mzo_std_sort_candidates (oracle) against the real std::sort with the
    reference's comparator, on inputs full of exact ties (where an unstable sort's result is algorithm-defined)"""
    import ctypes as C

    import numpy as np
    lib = hostsim_lib.load()
    orc = oracle_lib.load()
    orc.mzo_std_sort_candidates.argtypes = [C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    rng = np.random.default_rng(0)
    cases = []
    for n in list(range(0, 40)) + [64, 65, 82, 100, 200, 362]:
        for levels in (1, 2, 3, 7, 50, 10 ** 6):
            for _ in range(6):
                cases.append(rng.integers(0, levels, size=n).astype(np.float32) / np.float32(levels))
    for n in (82, 362):  # shapes that push the quicksort towards its depth limit
        cases.append(np.arange(n, dtype=np.float32))
        cases.append(np.arange(n, dtype=np.float32)[::-1].copy())
        cases.append(np.concatenate([np.arange(n // 2), np.arange(n - n // 2)[::-1]]).astype(np.float32))
        cases.append(np.zeros(n, np.float32))
        k = np.arange(n)
        cases.append(np.where(k % 2 == 0, k, n - k).astype(np.float32))
    for pol in cases:
        n = pol.size
        order = np.zeros(max(n, 1), np.int32)
        assert lib.hs_sort_matches_std(n, pol.ctypes.data_as(C.POINTER(C.c_float)), order.ctypes.data_as(C.POINTER(C.c_int32))) == 1, (n, pol[:20])
        a, p, l = np.arange(n, dtype=np.int32), pol.copy(), np.zeros(n, np.float32)
        orc.mzo_std_sort_candidates(n, a.ctypes.data_as(C.POINTER(C.c_int32)), p.ctypes.data_as(C.POINTER(C.c_float)), l.ctypes.data_as(C.POINTER(C.c_float)))
        assert np.array_equal(a, order[:n]), (n, pol[:20])


@pytest.mark.parametrize("name,game,n", [("env_ttt", 0, 3), ("env_go5", 1, 5), ("env_go9", 1, 9), ("env_go19", 1, 19), ("env_othello8", 2, 8), ("env_nogo9", 3, 9), ("env_gomoku15", 4, 15), ("env_hex11", 5, 11), ("env_killallgo7", 7, 7)])
def test_search_core_env_matches_reference_playouts(name, game, n):
    import env_replay
    case = env_replay.load(name)
    eng = hostsim_lib.HostSimSearch(hostsim_lib.load(), game, n, 1, 1)
    assert env_replay.replay(eng, case, check_score=lambda e: e.last_score) == case["game"].size


THINK_CASES = {"think_ttt_s50_k4": (0, 3), "think_go5_s60_k8": (1, 5), "think_go9_s100_k16_det": (1, 9), "think_go5_s23_k5": (1, 5),
               "think_othello_mz_s30_k6": (2, 8), "think_go5_mz_s20_k4": (1, 5)}


@pytest.mark.parametrize("name", list(THINK_CASES))
def test_search_core_think_matches_reference_recording(name):
    """console search (ZeroActor::think with a selection batch): the lane views, virtual loss in selection, duplicates, slot bookkeeping"""
    game, n = THINK_CASES[name]
    case = golden_replay.load_case(name)
    eng = hostsim_lib.HostSimSearch(hostsim_lib.load(), game, n, 1, int(case["S"]), think_k=int(case["K"]), **oracle_lib.conf_overrides(case["conf"]))
    assert golden_replay.replay_think(eng, case, features_of_duplicates=False) == case["move_action"].size
