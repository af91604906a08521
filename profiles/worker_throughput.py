"""Games per second through the drop-in worker executable itself (minizero_b200/bin/mz_sp, driven over stdin / stdout exactly
as the zero server drives `-mode sp`) on BASELINE configs[1]: Go 9x9, 400 simulations, 256 games per GPU, 6b x 256, training-default
stochasticity (Dirichlet noise, random rotations, softmax-count move choice, resign with the default ratio).

    python profiles/worker_throughput.py [seconds] [num_gpus]
"""
import json
import os
import re
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "minizero_b200", "bin", "mz_sp")
NET = os.path.join(ROOT, "oracle", "_ref", "nets", "go9_az_6bx256.pt")


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    gpus = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    games = 256 * gpus
    conf = (f"env_board_size=9:actor_num_simulation=400:zero_num_parallel_games={games}:nn_type_name=alphazero:nn_file_name={NET}:"
            "program_seed=1:program_auto_seed=false:program_quiet=true")
    env = dict(os.environ, CUDA_VISIBLE_DEVICES=",".join(str(i) for i in range(gpus)))
    p = subprocess.Popen([BIN, "-mode", "sp", "-conf_str", conf], stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
    lines = []

    def reader():
        for line in p.stdout:
            lines.append((time.perf_counter(), line))

    t = threading.Thread(target=reader, daemon=True)
    t.start()
    errs = []
    threading.Thread(target=lambda: errs.extend(p.stderr), daemon=True).start()
    p.stdin.write("start\n")
    p.stdin.flush()
    t0 = time.perf_counter()
    time.sleep(seconds)
    p.stdin.write("quit\n")
    p.stdin.flush()
    try:
        p.wait(timeout=60)
    except subprocess.TimeoutExpired:
        p.kill()
    t1 = time.perf_counter()
    got = [(ts, l) for ts, l in lines if l.startswith("SelfPlay ")]
    moves = sum(int(l.split()[3]) for _, l in got)
    # steady state: from the first finished game to the last one
    span = (got[-1][0] - got[0][0]) if len(got) > 1 else float("nan")
    out = {"workload": "go9x9_alphazero_400sims_256games_per_gpu_6bx256 through mz_sp (wire protocol)", "n_gpus": gpus, "wall_s": t1 - t0, "games": len(got),
           "moves_in_finished_games": moves, "mean_game_length": moves / max(1, len(got)), "games_per_sec_wall": len(got) / (t1 - t0),
           "games_per_sec_first_to_last_line": (len(got) - 1) / span if len(got) > 1 else None,
           "leaf_evals_per_sec_from_finished_games": moves * 401 / (t1 - t0), "all_lines_are_selfplay": len(got) == len(lines),
           "worker_timing": next((e.strip() for e in errs if e.startswith("[timing]")), None)}
    # every engine move is one whole search of that engine's 256 games (the [timing] line counts them): the steady-state rate of the executable
    m = re.search(r"\[timing\] (\d+) moves", out["worker_timing"] or "")
    if m:
        out["engine_moves"] = int(m.group(1))
        out["leaf_evals_per_sec_wall_incl_startup"] = int(m.group(1)) * 256 * 401 / (t1 - t0)  # model load, NCCL set-up and graph capture are inside the wall time
        # steady state from the worker's own clock: the phases of a move add up to the host thread's time per engine move, every engine runs its own loop
        ms = sum(float(x) for x in re.findall(r"(?:draw|tables|records|play|restart) ([0-9.]+)", out["worker_timing"]))
        out["ms_per_engine_move"] = ms
        out["leaf_evals_per_sec_steady"] = gpus * 256 * 401 / ms * 1e3
    print(json.dumps(out))


if __name__ == "__main__":
    main()
