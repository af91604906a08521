"""Generate environment differential-playout fixtures (tests/golden/env_*.npz) with the UNMODIFIED reference environments
(oracle/_ref/ref_env_playout_*): random legal playouts, with the legal set, rotated feature planes, terminal flag and score
at every step. Only runs in the build container. TEST INFRASTRUCTURE ONLY."""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden")
# name: (binary, conf, seed, games, max moves per game)
CASES = {
    "env_ttt": ("tictactoe", "program_quiet=true", 1, 24, 20),
    "env_go5": ("go", "env_board_size=5:program_quiet=true", 2, 8, 100),
    "env_go9": ("go", "env_board_size=9:program_quiet=true", 3, 4, 400),
    "env_go9_situational": ("go", "env_board_size=9:env_go_ko_rule=situational:program_quiet=true", 4, 2, 400),
    "env_go19": ("go", "env_board_size=19:program_quiet=true", 5, 1, 420),
    "env_nogo9": ("nogo", "program_quiet=true", 7, 6, 200),
    "env_gomoku15": ("gomoku", "program_quiet=true", 8, 5, 230),
    "env_gomoku15_freestyle": ("gomoku", "env_gomoku_exactly_five_stones=false:env_gomoku_rule=outer_open:program_quiet=true", 9, 3, 230),
    "env_hex11": ("hex", "program_quiet=true", 10, 10, 130),
    "env_hex11_noswap": ("hex", "env_hex_use_swap_rule=false:program_quiet=true", 11, 4, 130),
    "env_othello8": ("othello", "program_quiet=true", 6, 8, 200),
    # KillAllGo 7x7 (seki table off, the default): the opening rule, Benson's unconditional life as the terminal test, the result
    "env_killallgo7": ("killallgo", "program_quiet=true", 13, 60, 120),
}


def main(names):
    for name in names:
        binary, conf, seed, games, max_moves = CASES[name]
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, "p.bin")
            # KillAllGo's set-up loads 7x7_seki.db from the working directory and otherwise GENERATES it with a search of its own (killallgo.cpp:10-25):
            # an empty table (a zero entry count) stands in; env_killallgo_use_seki is false, the table is never consulted
            with open(os.path.join(d, "7x7_seki.db"), "wb") as f:
                f.write((0).to_bytes(8, "little"))
            res = subprocess.run([os.path.join(HERE, "_ref", "ref_env_playout_" + binary), conf, str(seed), str(games), str(max_moves), path], check=True,
                                 capture_output=True, text=True, cwd=d)
            tok = res.stdout.split()
            A, F = int(tok[1]), int(tok[3])
            raw = np.fromfile(path, dtype=np.uint8).reshape(-1, 28 + A + F)
            hdr = raw[:, :24].copy().view(np.int32)
            score = raw[:, 24:28].copy().view(np.float32)[:, 0]
            np.savez_compressed(os.path.join(OUT, name + ".npz"), A=A, F=F, conf=conf, game=hdr[:, 0], step=hdr[:, 1], turn=hdr[:, 2], rotation=hdr[:, 3].astype(np.uint8),
                                action=hdr[:, 4], terminal_after=hdr[:, 5], score_after=score, legal=np.packbits(raw[:, 28:28 + A], axis=1),
                                features=np.packbits(raw[:, 28 + A:], axis=1))
            print(name, "steps", raw.shape[0], "terminal games", int(hdr[:, 5].sum()), os.path.getsize(os.path.join(OUT, name + ".npz")) // 1024, "KiB")


if __name__ == "__main__":
    main(sys.argv[1:] or list(CASES))
