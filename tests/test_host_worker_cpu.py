"""Host side of the drop-in worker (minizero_b200/host), CPU-only checks:
  * the SelfPlay line / game record formatter reproduces, byte for byte, every line the compiled reference printed in the
    golden recordings (terminal and resigned games);
  * the configuration loader accepts the reference's keys and refuses unknown ones."""
import os
import re
import struct
import subprocess

import numpy as np
import pytest

import golden_replay
import oracle_lib

ROOT = oracle_lib.ROOT
BIN = os.path.join(ROOT, "minizero_b200", "bin", "mz_sp")


def hexf(x):
    return "%08x" % struct.unpack("<I", struct.pack("<f", float(x)))[0]


@pytest.fixture(scope="module")
def worker_binary():
    if not os.path.exists(BIN):
        import __graft_entry__ as ge
        ge.build()
    return BIN


@pytest.mark.parametrize("name,gname,board,komi", [("ttt_s50_b2", "tictactoe", 3, None), ("ttt_s50_b1_det", "tictactoe", 3, None), ("go5_s24_b2", "go_5x5", 5, 7.5)])
def test_selfplay_lines_match_reference_bytes(worker_binary, name, gname, board, komi):
    z = golden_replay.load_case(name)
    lines = [str(l) for l in z["selfplay_lines"]]
    games = []
    for g in range(int(z["B"])):
        cur = []
        for m in [m for m in range(z["move_game"].size) if z["move_game"][m] == g]:
            if cur and int(z["move_number"][m]) == 0:
                games.append(cur)
                cur = []
            cur.append(m)
            if z["move_resign"][m]:
                games.append(cur)
                cur = []
        if cur:
            games.append(cur)
    produced = set()
    for gm in games:
        resigned = bool(z["move_resign"][gm[-1]])
        moves = gm[:-1] if resigned else gm
        turn = 1 if len(moves) % 2 == 0 else 2
        for eval_score in (1.0, -1.0, 0.0):
            inp = [f"header {gname} {board} {0 if komi is None else 1} {komi or 0} /some/dir/{name_of_model(lines)} {0 if resigned else 1} {hexf(eval_score)} {turn}"]
            for m in moves:
                k = int(z["move_num_children"][m])
                pairs = " ".join(f"{int(z['child_action'][m, i])}:{hexf(z['child_count'][m, i])}" for i in range(k))
                inp.append(f"move {int(z['move_player'][m])} {int(z['move_action'][m])} {hexf(z['root_mean'][m])} {k} {pairs}")
            r = subprocess.run([worker_binary, "-mode", "record_test"], input="\n".join(inp) + "\n", capture_output=True, text=True, check=True)
            produced.add(r.stdout.strip())
    for l in lines:
        assert l.strip() in produced, l[:120]


def name_of_model(lines):
    return re.search(r"EV\[([^\]]*)\]", lines[0]).group(1)


def test_config_accepts_reference_keys_and_refuses_unknown(worker_binary):
    ok = subprocess.run([worker_binary, "-mode", "record_test", "-conf_str", "actor_num_simulation=5"], input="", capture_output=True, text=True)
    assert ok.returncode == 0
    bad = subprocess.run([worker_binary, "-mode", "sp", "-conf_str", "no_such_key=1"], input="", capture_output=True, text=True)
    assert bad.returncode != 0 and "Invalid key" in bad.stderr
    # without a GPU (or a model) the worker must fail loudly and write nothing to stdout
    nogpu = subprocess.run([worker_binary, "-mode", "sp", "-conf_str", "nn_file_name=/nonexistent.pt:zero_num_parallel_games=2"], input="quit\n", capture_output=True, text=True)
    assert nogpu.returncode != 0 and nogpu.stdout == ""


def test_gumbel_selfplay_lines_match_reference_bytes(worker_binary):
    """Othello Gumbel MuZero records: the completed-Q policy tags (gumbel_zero.cpp:9-59, unordered_map order included) and the rest
    of the line, byte for byte as the compiled reference printed them"""
    z = golden_replay.load_case("othello_gmz_s16_b2")
    lines = [str(l) for l in z["selfplay_lines"]]
    assert lines
    produced = set()
    for g in range(int(z["B"])):
        ms = [m for m in range(z["move_game"].size) if z["move_game"][m] == g]
        games, cur = [], []
        for m in ms:
            if cur and int(z["move_number"][m]) == 0:
                games.append(cur)
                cur = []
            cur.append(m)
        games.append(cur)
        for gm in games:
            for eval_score in (1.0, -1.0, 0.0):
                inp = [f"header othello_8x8 8 0 0 /some/dir/{name_of_model(lines)} 1 {hexf(eval_score)} 1"]
                for m in gm:
                    k = int(z["move_num_children"][m])
                    toks = " ".join(":".join([str(int(z["child_action"][m, i]))] + [hexf(z["child_" + n][m, i]) for n in ("count", "mean", "policy", "logit", "noise")])
                                    for i in range(k))
                    inp.append(f"gmove {int(z['move_player'][m])} {int(z['move_action'][m])} {hexf(z['root_mean'][m])} {hexf(z['root_value'][m])} {int(z['S'])} 50 1 {k} {toks}")
                r = subprocess.run([worker_binary, "-mode", "record_test"], input="\n".join(inp) + "\n", capture_output=True, text=True, check=True)
                produced.add(r.stdout.strip())
    for l in lines:
        assert l.strip() in produced, l[:200]


def test_intermediate_sequence_lines_match_reference_bytes(worker_binary):
    """zero_actor_intermediate_sequence_length > 0 (actor_group.cpp:24-64,129-131): when a piece is due, its data range, the
    `false` terminal flag, the resign-style return, and the action info dropped from moves already sent"""
    z = golden_replay.load_case("go5_seq_s8_b2")
    lines = [str(l) for l in z["selfplay_lines"]]
    assert len(lines) >= 8 and any(l.startswith("SelfPlay false") for l in lines)
    produced = set()
    for g in range(int(z["B"])):
        games, cur = [], []
        for m in [m for m in range(z["move_game"].size) if z["move_game"][m] == g]:
            if cur and int(z["move_number"][m]) == 0:
                games.append(cur)
                cur = []
            cur.append(m)
        games.append(cur)
        for gm in games:
            resigned = bool(z["move_resign"][gm[-1]])  # the resigning search's move is not played; the game is sent as a non-terminal piece
            gm = gm[:-1] if resigned else gm
            turn = 1 if len(gm) % 2 == 0 else 2
            for eval_score in (1.0, -1.0, 0.0):
                inp = [f"header go_5x5 5 1 7.5 /some/dir/{name_of_model(lines)} {0 if resigned else 1} {hexf(eval_score)} {turn} 8 2 3"]
                for m in gm:
                    k = int(z["move_num_children"][m])
                    pairs = " ".join(f"{int(z['child_action'][m, i])}:{hexf(z['child_count'][m, i])}" for i in range(k))
                    inp.append(f"move {int(z['move_player'][m])} {int(z['move_action'][m])} {hexf(z['root_mean'][m])} {k} {pairs}")
                r = subprocess.run([worker_binary, "-mode", "record_test"], input="\n".join(inp) + "\n", capture_output=True, text=True, check=True)
                produced.update(x.strip() for x in r.stdout.splitlines())
    for l in lines:
        assert l.strip() in produced, l[:160]


@pytest.mark.parametrize("name,game_type,board", [("go5_s24_b2", 1, 5), ("ttt_s50_b2", 0, 3), ("go9_s32_b2", 1, 9), ("othello_gmz_s16_b2", 2, 8),
                                                  ("othello_mz_s24_b2", 2, 8), ("nogo9_s8_b2", 3, 9), ("gomoku15_s8_b2", 4, 15), ("hex11_s8_b2", 5, 11), ("killallgo7_s16_b2", 7, 7)])
def test_host_draw_sequence_matches_reference_seed(worker_binary, name, game_type, board):
    """Seed-exact randomness: fed with the root tables the reference saw, the worker's host logic — the same
    member functions the GPU path uses, libstdc++'s mt19937 and distributions in the reference's order — reproduces every rotation
    the reference drew, its Dirichlet noise bit for bit, its softmax-count move choices and its resignations"""
    z = golden_replay.load_case(name)
    B, S, A = int(z["B"]), int(z["S"]), int(z["A"])
    conf = str(z["conf"])
    per_game = {g: [m for m in range(z["move_game"].size) if z["move_game"][m] == g] for g in range(B)}
    rounds = min(len(v) for v in per_game.values())
    inp = [f"setup {game_type} {board} {A}"]
    for r in range(rounds):
        for g in range(B):
            m = per_game[g][r]
            k = int(z["move_num_children"][m])
            toks = " ".join(":".join([str(int(z["child_action"][m, i]))] + [hexf(z["child_" + n][m, i]) for n in ("count", "mean", "policy", "logit", "noise")])
                            for i in range(k))
            inp.append(f"root {g} {k} {hexf(z['root_mean'][m])} {hexf(z['root_value'][m])} {toks}")
    out = subprocess.run([worker_binary, "-mode", "rng_test", "-conf_str", conf], input="\n".join(inp) + "\n", capture_output=True, text=True, check=True).stdout
    rot, noise, acts = {}, [], []
    for line in out.splitlines():
        f = line.split()
        if f[0] == "rot":
            rot[(int(f[1]), int(f[2]))] = int(f[3])
        elif f[0] == "noise":
            noise.append((int(f[1]), [int(x, 16) for x in f[2:]]))
        elif f[0] == "act":
            acts.append((int(f[1]), int(f[2]), int(f[3])))
    n_cycles = z["eval_game"].size // B
    checked = 0
    for c in range(min(n_cycles, rounds * (S + 1))):
        for g in range(B):
            assert rot[(c, g)] == int(z["eval_rotation"][c * B + g]), (c, g)
            checked += 1
    assert checked >= rounds * (S + 1) * B - B
    assert len(noise) == rounds * B and len(acts) == rounds * B
    for r in range(rounds):
        for g in range(B):
            m = per_game[g][r]
            k = int(z["move_num_children"][m])
            gg, bits = noise[r * B + g]
            assert gg == g and bits == z["child_noise"][m, :k].view(np.uint32).tolist(), (r, g)
            gg, action, resign = acts[r * B + g]
            assert gg == g and resign == int(z["move_resign"][m]), (r, g)
            assert action == (-1 if resign else int(z["move_action"][m])), (r, g, action, int(z["move_action"][m]))  # a resigning search plays nothing


@pytest.mark.parametrize("name", ["atari_mz_s20_b2", "atari_mz_s50_b2_det", "atari_mz_s18_gumbel_b2", "atari_mz_seq_s8_b2"])
def test_atari_records_and_draws_match_reference(worker_binary, name):
    """Atari mode of the worker, host side end to end without a device: fed with the root tables (children's rewards and value bounds
    included) the -DATARI reference saw, the worker draws the emulator seeds, root noise, moves and resignations of the reference's
    run, steps the synthetic frame source itself, and prints every SelfPlay line — OBS (gzip + hex of the recent screens), SD, L tags,
    rewards, intermediate sequences — byte for byte as the compiled reference printed it"""
    z = golden_replay.load_case(name)
    B, S, A = int(z["B"]), int(z["S"]), int(z["A"])
    conf = str(z["conf"]) + ":nn_file_name=/some/dir/atari_mz_1bx32.pt"
    per_game = {g: [m for m in range(z["move_game"].size) if z["move_game"][m] == g] for g in range(B)}
    rounds = min(len(v) for v in per_game.values())
    inp = ["setup 6 6 %d" % A]
    for r in range(rounds):
        for g in range(B):
            m = per_game[g][r]
            k = int(z["move_num_children"][m])
            toks = " ".join(":".join([str(int(z["child_action"][m, i]))] + [hexf(z["child_" + n][m, i]) for n in ("count", "mean", "policy", "logit", "noise", "reward")])
                            for i in range(k))
            inp.append(f"bounds {int(z['bound_size'][m])} {hexf(z['bound_lo'][m])} {hexf(z['bound_hi'][m])}")
            inp.append(f"root {g} {k} {hexf(z['root_mean'][m])} {hexf(z['root_value'][m])} {toks}")
    out = subprocess.run([worker_binary, "-mode", "rng_test", "-conf_str", conf], input="\n".join(inp) + "\n", capture_output=True, text=True, check=True).stdout
    noise, acts, lines = [], [], []
    for line in out.splitlines():
        f = line.split(" ", 1)
        if f[0] == "noise":
            noise.append([int(x, 16) for x in f[1].split()[1:]])
        elif f[0] == "act":
            acts.append([int(x) for x in f[1].split()])
        elif f[0] == "line":
            lines.append(f[1])
    assert len(acts) == rounds * B
    for r in range(rounds):
        for g in range(B):
            m = per_game[g][r]
            k = int(z["move_num_children"][m])
            assert noise[r * B + g] == z["child_noise"][m, :k].view(np.uint32).tolist(), (r, g)
            gg, action, resign, ended = acts[r * B + g]
            assert gg == g and resign == int(z["move_resign"][m]) and action == int(z["move_action"][m]), (r, g)
            assert ended == int(z["env_terminal"][m]), (r, g)  # the worker's emulator ends the episode where the reference's did
    want = [str(l) for l in z["selfplay_lines"]]
    assert len(lines) >= len(want) - 1 and (want or name == "atari_mz_s50_b2_det")  # (the short deterministic recording finishes no episode)
    for i, l in enumerate(lines[:len(want)]):
        assert l in want, (i, l[:100])
    assert sum(1 for l in want if l in lines) >= len(want) - 1
