"""The drop-in worker executable (minizero_b200/bin/mz_sp) end to end on a B200: it is driven over stdin exactly as the zero
server drives the reference's `-mode sp` process, and every `SelfPlay` line it prints is handed to the REFERENCE's own record
loader and rules (oracle/_ref/ref_record_check_*), which must parse it, replay every move as legal and agree on result,
lengths and return."""
import os
import subprocess
import tempfile
import time

import pytest

import oracle_lib

pytestmark = pytest.mark.gpu
ROOT = oracle_lib.ROOT
BIN = os.path.join(ROOT, "minizero_b200", "bin", "mz_sp")
NETS = os.path.join(ROOT, "oracle", "_ref", "nets")


def run_worker(conf, want_lines, timeout=240):
    p = subprocess.Popen([BIN, "-mode", "sp", "-conf_str", conf], stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    p.stdin.write("keep_alive\nstart\n")
    p.stdin.flush()
    lines, t0 = [], time.time()
    while len(lines) < want_lines and time.time() - t0 < timeout:
        line = p.stdout.readline()
        if not line:
            break
        lines.append(line.rstrip("\n"))
    p.stdin.write("quit\n")
    p.stdin.flush()
    try:
        out, err = p.communicate(timeout=60)
    except subprocess.TimeoutExpired:
        p.kill()
        out, err = p.communicate()
    lines += [l for l in out.splitlines() if l]
    return lines, err, p.returncode


@pytest.mark.parametrize("game,net,conf,checker_conf", [
    ("tictactoe", "ttt_az_2bx32", "actor_num_simulation=50:zero_num_parallel_games=16", ""),
    ("go", "go5_az_1bx16", "env_board_size=5:actor_num_simulation=24:zero_num_parallel_games=16", "env_board_size=5"),
    ("go", "go9_az_2bx64", "env_board_size=9:actor_num_simulation=32:zero_num_parallel_games=32", "env_board_size=9"),
    # BASELINE configs[2] search settings (tools/quick-run.sh "gmz"): Gumbel MuZero on Othello
    ("othello", "othello_mz_1bx32", "actor_num_simulation=16:zero_num_parallel_games=16:nn_type_name=muzero:actor_use_gumbel=true:actor_use_gumbel_noise=true:"
     "actor_gumbel_sample_size=16:actor_use_dirichlet_noise=false", ""),
    ("othello", "othello_mz_1bx32", "actor_num_simulation=24:zero_num_parallel_games=16:nn_type_name=muzero", ""),
    ("go", "go5_mz_1bx16", "env_board_size=5:actor_num_simulation=16:zero_num_parallel_games=16:nn_type_name=muzero", "env_board_size=5"),
    ("nogo", "nogo9_az_1bx16", "actor_num_simulation=16:zero_num_parallel_games=16", ""),
    ("gomoku", "gomoku15_az_1bx16", "actor_num_simulation=8:zero_num_parallel_games=24", ""),
    ("hex", "hex11_az_1bx16", "actor_num_simulation=8:zero_num_parallel_games=24", ""),
    # KillAllGo 7x7: the worker ends games by Benson's unconditional life on its own copy of the position, in the reference's draw order
    ("killallgo", "killallgo7_az_1bx16", "actor_num_simulation=16:zero_num_parallel_games=16", ""),
])
def test_worker_speaks_the_wire_protocol_and_reference_accepts_its_records(game, net, conf, checker_conf):
    checker = os.path.join(ROOT, "oracle", "_ref", "ref_record_check_" + game)
    model = os.path.join(NETS, net + ".pt")
    if not (os.path.exists(BIN) and os.path.exists(checker) and os.path.exists(model)):
        pytest.skip("worker binary / oracle/_ref not built")
    conf = conf + f":nn_file_name={model}:program_seed=3:program_auto_seed=false:program_quiet=true:zero_num_threads=1"
    lines, err, rc = run_worker(conf, want_lines=24)
    assert rc == 0, err[-500:]
    assert len(lines) >= 24, err[-500:]
    assert all(l.startswith("SelfPlay ") and l.endswith(" #") for l in lines), "stdout must carry nothing but SelfPlay lines (zero_server.cpp:130-139)"
    with tempfile.TemporaryDirectory() as d:  # the KillAllGo build of the reference wants its seki table in the working directory: an empty one (unused)
        with open(os.path.join(d, "7x7_seki.db"), "wb") as f:
            f.write((0).to_bytes(8, "little"))
        r = subprocess.run([checker, checker_conf], input="\n".join(lines) + "\n", capture_output=True, text=True, cwd=d)
    assert r.stdout.strip() == f"RECORDS_OK {len(lines)}", r.stdout + r.stderr[-300:]
    assert f"EV[{net}.pt]" in lines[0]


def test_worker_reloads_model_and_updates_config():
    model = os.path.join(NETS, "ttt_az_2bx32.pt")
    if not (os.path.exists(BIN) and os.path.exists(model)):
        pytest.skip("worker binary / nets not built")
    conf = f"actor_num_simulation=20:zero_num_parallel_games=8:nn_file_name={model}:program_quiet=true"
    p = subprocess.Popen([BIN, "-mode", "sp", "-conf_str", conf], stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    p.stdin.write(f"load_model {model}\nupdate_config actor_num_simulation=20:actor_select_action_by_count=true:actor_select_action_by_softmax_count=false\nreset_actors\nstart\n")
    p.stdin.flush()
    first = p.stdout.readline()
    p.stdin.write("stop\nquit\n")
    p.stdin.flush()
    out, err = p.communicate(timeout=120)
    assert first.startswith("SelfPlay ") and p.returncode == 0, err[-400:]
    assert "[ignored command] reset_actors" in err  # zero_actor_ignored_command default (configuration.cpp:47)


def test_worker_atari_mode_end_to_end():
    """BASELINE configs[4] through the executable: Atari MuZero with value rescaling on the synthetic frame source; every line must
    carry an OBS tag that gunzips to whole 96 x 96 x 3 screens covering the sequence's window, an SD seed, rewards summing to the return"""
    import gzip
    import re
    model = os.path.join(NETS, "atari_mz_1bx32.pt")
    if not (os.path.exists(BIN) and os.path.exists(model)):
        pytest.skip("worker binary / nets not built")
    conf = ("env_atari_name=ms_pacman:actor_mcts_value_rescale=true:actor_mcts_reward_discount=0.997:actor_num_simulation=20:zero_num_parallel_games=16:"
            f"nn_type_name=muzero:zero_actor_intermediate_sequence_length=8:learner_n_step_return=3:learner_muzero_unrolling_step=2:nn_file_name={model}:"
            "program_seed=3:program_auto_seed=false:program_quiet=true:zero_num_threads=1")
    lines, err, rc = run_worker(conf, want_lines=24)
    assert rc == 0, err[-500:]
    assert len(lines) >= 24 and all(l.startswith("SelfPlay ") and l.endswith(" #") for l in lines)
    terminal_seen = False
    for l in lines:
        f = l.split(" ")
        terminal, data_len, game_len, ret = f[1] == "true", int(f[2]), int(f[3]), float(f[4])
        rec = f[5]
        assert rec.startswith("(;GM[atari_ms_pacman]RE[") and "EV[atari_mz_1bx32.pt]" in rec and re.search(r"SD\[\d+\]", rec)
        obs = gzip.decompress(bytes.fromhex(re.search(r"OBS\[([0-9a-f]*)\]", rec).group(1)))
        assert len(obs) % (3 * 96 * 96) == 0 and len(obs) // (3 * 96 * 96) == min(game_len + 1, 8 + 8 + 3 + 2 + 1)
        moves = re.findall(r";B\[(\d+)\]([^;)]*)", rec)
        assert len(moves) == game_len and 1 <= data_len <= game_len
        assert all(int(a) in (0, 2, 3, 4, 5, 6, 7, 8, 9) for a, _ in moves)  # the minimal action set of the synthetic game
        rewards = [float(re.search(r"R\[([^\]]*)\]", info).group(1)) for _, info in moves if "R[" in info]
        if len(rewards) == game_len:  # no action info dropped yet: the rewards add up to the return
            assert sum(rewards) == ret
        terminal_seen |= terminal
    assert terminal_seen


def test_worker_on_all_visible_gpus_one_host_thread_per_engine():
    """more than one GPU: the engines are driven by their own host threads (zero_num_threads > 1, Worker::runThreaded), the weights reach every
    GPU through the NCCL broadcast, and every record of every engine is accepted by the reference's loader and rules"""
    torch = pytest.importorskip("torch")
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    checker = os.path.join(ROOT, "oracle", "_ref", "ref_record_check_go")
    model = os.path.join(NETS, "go9_az_2bx64.pt")
    if not (os.path.exists(BIN) and os.path.exists(checker) and os.path.exists(model)):
        pytest.skip("worker binary / oracle/_ref not built")
    conf = (f"env_board_size=9:actor_num_simulation=32:zero_num_parallel_games={16 * n}:nn_file_name={model}:program_seed=3:program_auto_seed=false:"
            "program_quiet=true:zero_num_threads=4")
    lines, err, rc = run_worker(conf, want_lines=16 * n + 8)
    assert rc == 0, err[-500:]
    assert len(lines) >= 16 * n + 8 and all(l.startswith("SelfPlay ") and l.endswith(" #") for l in lines)
    r = subprocess.run([checker, "env_board_size=9"], input="\n".join(lines) + "\n", capture_output=True, text=True)
    assert r.stdout.strip() == f"RECORDS_OK {len(lines)}", r.stdout + r.stderr[-300:]


@pytest.mark.parametrize("game,net,conf,checker_conf", [
    ("go", "go9_az_2bx64", "env_board_size=9:actor_num_simulation=32:zero_num_parallel_games=32", "env_board_size=9"),
    ("tictactoe", "ttt_az_2bx32", "actor_num_simulation=50:zero_num_parallel_games=16", ""),
    ("othello", "othello_mz_1bx32", "actor_num_simulation=24:zero_num_parallel_games=16:nn_type_name=muzero", ""),
])
def test_reference_actor_group_bound_to_the_library(game, net, conf, checker_conf):
    """integration/b200_actor_group.cpp: the REFERENCE's ActorGroup, actors, move decision and record writer (compiled from the unmodified sources)
    with every search run by libmzb200 through the C ABI; driven over the wire protocol, records checked by the reference's loader"""
    binding = os.path.join(ROOT, "oracle", "_ref", "b200_actor_group_" + game)
    checker = os.path.join(ROOT, "oracle", "_ref", "ref_record_check_" + game)
    model = os.path.join(NETS, net + ".pt")
    if not (os.path.exists(binding) and os.path.exists(checker) and os.path.exists(model)):
        pytest.skip("binding / oracle/_ref not built (needs the reference checkout)")
    conf = conf + f":nn_file_name={model}:program_seed=5:program_auto_seed=false:program_quiet=true:zero_num_threads=2"
    p = subprocess.Popen([binding, conf], stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    p.stdin.write("start\n")
    p.stdin.flush()
    lines, t0 = [], time.time()
    while len(lines) < 20 and time.time() - t0 < 240:
        line = p.stdout.readline()
        if not line:
            break
        lines.append(line.rstrip("\n"))
    p.stdin.write("quit\n")
    p.stdin.flush()
    try:
        p.communicate(timeout=30)
    except subprocess.TimeoutExpired:  # the reference's quit is exit(0) from the command thread: some builds linger in thread teardown
        p.kill()
    assert len(lines) >= 20 and all(l.startswith("SelfPlay ") and l.endswith(" #") for l in lines)
    r = subprocess.run([checker, checker_conf], input="\n".join(lines) + "\n", capture_output=True, text=True)
    assert r.stdout.strip() == f"RECORDS_OK {len(lines)}", r.stdout + r.stderr[-300:]
    assert f"EV[{net}.pt]" in lines[0]
