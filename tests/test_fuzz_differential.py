"""CPU: the oracle (oracle/port, plain C) against the host-sim build of the CUDA source (search_core.cuh compiled with a 1-lane warp),
driven with the same synthetic network outputs. See fuzz_differential.py."""
import zlib

import pytest

import fuzz_differential
import hostsim_lib
import oracle_lib

CASES = [
    # name, game, board, A, sims, games, moves, options
    ("go5_az", 1, 5, 26, 200, 3, 40, {}),
    ("go9_az", 1, 9, 82, 120, 2, 30, {}),
    ("ttt_az", 0, 3, 9, 100, 3, 14, {}),
    ("nogo9_az", 3, 9, 82, 60, 2, 90, {}),
    ("othello_az", 2, 8, 65, 80, 2, 70, {}),
    ("gomoku15_az", 4, 15, 225, 40, 2, 120, {}),
    ("hex11_az", 5, 11, 121, 40, 2, 130, {}),
    ("othello_muzero", 2, 8, 65, 64, 2, 40, dict(muzero=1)),
    ("othello_gumbel_muzero_m8", 2, 8, 65, 32, 2, 66, dict(muzero=1, use_gumbel=1, gumbel_noise=1, gumbel_sample_size=8)),
    ("othello_gumbel_muzero_m16_s200", 2, 8, 65, 200, 2, 20, dict(muzero=1, use_gumbel=1, gumbel_noise=1, gumbel_sample_size=16)),
    ("go5_gumbel_muzero_m4", 1, 5, 26, 50, 2, 40, dict(muzero=1, use_gumbel=1, gumbel_noise=1, gumbel_sample_size=4)),
]


@pytest.mark.parametrize("name,game,board,A,sims,games,moves,opts", CASES, ids=[c[0] for c in CASES])
def test_oracle_and_search_core_agree_on_synthetic_searches(oracle, name, game, board, A, sims, games, moves, opts):
    orc = oracle_lib.OracleSearch(oracle, game, board, games, sims, **opts)
    sim = hostsim_lib.HostSimSearch(hostsim_lib.load(), game, board, games, sims, **opts)
    muzero, gumbel = bool(opts.get("muzero")), bool(opts.get("use_gumbel"))
    n = fuzz_differential.run(orc, sim, num_actions=A, sims=sims, games=games, moves=moves, seed=zlib.crc32(name.encode()) % 1000,
                              noise=("gumbel" if opts.get("gumbel_noise") else "dirichlet"), muzero=muzero, gumbel=gumbel)
    assert n == games * moves
