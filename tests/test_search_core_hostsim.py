"""search_core.cuh (the source of the CUDA tree/env kernels) compiled for the host with a 1-lane warp,
replayed against the recordings of the compiled reference. Catches logic errors without a GPU; the
-m gpu tests repeat the same replay through the real kernels and the C-ABI."""
import pytest

import golden_replay
import hostsim_lib

import oracle_lib

CASES = {"ttt_s50_b2": (0, 3), "ttt_s50_b1_det": (0, 3), "go5_s24_b2": (1, 5), "go9_s32_b2": (1, 9),
         "othello_gmz_s16_b2": (2, 8), "othello_gmz_s32_m8_b2": (2, 8), "othello_mz_s24_b2": (2, 8)}


@pytest.mark.parametrize("name", list(CASES))
def test_search_core_matches_reference_recording(name):
    game, n = CASES[name]
    case = golden_replay.load_case(name)
    eng = hostsim_lib.HostSimSearch(hostsim_lib.load(), game, n, int(case["B"]), int(case["S"]), **oracle_lib.conf_overrides(case["conf"]))
    checked = golden_replay.replay(eng, case)
    assert checked >= case["move_game"].size - int(case["B"])
