"""Generate golden vectors for tests/golden/ by running the UNMODIFIED reference (oracle/_ref/ref_stepper_*).

Only runs in the build container (needs oracle/_ref built from /root/reference and the nets from
oracle/gen_nets.py). Output: tests/golden/<case>.npz with, per leaf evaluation, the game index,
rotation, path length, bit-packed feature planes and the network outputs the reference consumed;
and per move the root child table at the moment the move was decided.
TEST INFRASTRUCTURE ONLY.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden")

COMMON = "zero_num_threads=1:program_seed=%d:program_auto_seed=false:program_quiet=true:nn_type_name=alphazero"
COMMON_MZ = "zero_num_threads=1:program_seed=%d:program_auto_seed=false:program_quiet=true:nn_type_name=muzero"
# BASELINE configs[2] search settings (tools/quick-run.sh:333-346 "gmz"), small net
GUMBEL = "actor_use_gumbel=true:actor_use_gumbel_noise=true:actor_gumbel_sample_size=%d:actor_gumbel_sigma_visit_c=50:actor_gumbel_sigma_scale_c=1:actor_use_dirichlet_noise=false:"
ATARI = "env_atari_name=ms_pacman:actor_mcts_value_rescale=true:actor_mcts_reward_discount=0.997:"
CASES = {
    # name: (binary, net, conf, max_moves)
    "ttt_s50_b2": ("tictactoe", "ttt_az_2bx32", "actor_num_simulation=50:zero_num_parallel_games=2:" + COMMON % 1, 40),
    "ttt_s50_b1_det": ("tictactoe", "ttt_az_2bx32", "actor_num_simulation=50:zero_num_parallel_games=1:actor_use_dirichlet_noise=false:actor_use_random_rotation_features=false:actor_select_action_by_count=true:actor_select_action_by_softmax_count=false:" + COMMON % 1, 12),
    "go5_s24_b2": ("go", "go5_az_1bx16", "env_board_size=5:actor_num_simulation=24:zero_num_parallel_games=2:" + COMMON % 3, 140),
    "go9_s32_b2": ("go", "go9_az_1bx16", "env_board_size=9:actor_num_simulation=32:zero_num_parallel_games=2:" + COMMON % 5, 30),
    # 19x19 (BASELINE configs[3] board): two-plane policy head, 362 actions, row bitboards at full width
    "go19_s8_b2": ("go", "go19_az_1bx16", "env_board_size=19:actor_num_simulation=8:zero_num_parallel_games=2:" + COMMON % 13, 30),
    # intermediate sequences (actor_group.cpp:24-64,129-131): long games sent in pieces, action info of sent moves dropped
    "go5_seq_s8_b2": ("go", "go5_az_1bx16", "env_board_size=5:actor_num_simulation=8:zero_num_parallel_games=2:zero_actor_intermediate_sequence_length=8:"
                      "learner_n_step_return=3:learner_muzero_unrolling_step=2:" + COMMON % 21, 110),
    # MuZero on the other board games (their action planes are the same one-hot cell: go.cpp:310-315, tictactoe.cpp:92-97)
    "go5_mz_s16_b2": ("go", "go5_mz_1bx16", "env_board_size=5:actor_num_simulation=16:zero_num_parallel_games=2:" + COMMON_MZ % 31, 90),
    "ttt_gmz_s16_b2": ("tictactoe", "ttt_mz_1bx16", "actor_num_simulation=16:zero_num_parallel_games=2:" + GUMBEL % 4 + COMMON_MZ % 32, 40),
    # NoGo 9x9 (environment/nogo/nogo.h): GoEnv with its own legality (no capture, no suicide, no pass), end and result
    "nogo9_s8_b2": ("nogo", "nogo9_az_1bx16", "actor_num_simulation=8:zero_num_parallel_games=2:" + COMMON % 41, 100),
    # Gomoku 15x15 (environment/gomoku): no pass, exactly five in a row through the last move wins
    "gomoku15_s8_b2": ("gomoku", "gomoku15_az_1bx16", "actor_num_simulation=8:zero_num_parallel_games=2:" + COMMON % 51, 80),
    # Hex 11x11 (environment/hex): swap rule, connect the two own edges; features / policy are not rotated although a rotation is drawn
    "hex11_s8_b2": ("hex", "hex11_az_1bx16", "actor_num_simulation=8:zero_num_parallel_games=2:" + COMMON % 61, 110),
    # KillAllGo 7x7 (environment/killallgo, seki table off): Black's two-stone opening, Benson's unconditional life ends the game inside the tree and at the root
    "killallgo7_s16_b2": ("killallgo", "killallgo7_az_1bx16", "actor_num_simulation=16:zero_num_parallel_games=2:" + COMMON % 95, 120),
    # Othello 8x8 MuZero: Gumbel (configs[2] settings: n=16, m=16), Gumbel with real halving (n=32, m=8), plain PUCT MuZero with Dirichlet noise
    "othello_gmz_s16_b2": ("othello", "othello_mz_1bx32", "actor_num_simulation=16:zero_num_parallel_games=2:" + GUMBEL % 16 + COMMON_MZ % 7, 130),
    "othello_gmz_s32_m8_b2": ("othello", "othello_mz_1bx32", "actor_num_simulation=32:zero_num_parallel_games=2:" + GUMBEL % 8 + COMMON_MZ % 8, 70),
    "othello_mz_s24_b2": ("othello", "othello_mz_1bx32", "actor_num_simulation=24:zero_num_parallel_games=2:" + COMMON_MZ % 9, 70),
    # Gumbel sample sizes that are not powers of two: the halving divisor log2(m) * sample_size / 2 is evaluated in double
    # (gumbel_zero.cpp:109), so 12 -> 6 -> 3 -> 1 divides by 1.5 * log2(m) at sample size 3
    "go5_gmz_s64_m12_b2": ("go", "go5_mz_1bx16", "env_board_size=5:actor_num_simulation=64:zero_num_parallel_games=2:" + GUMBEL % 12 + COMMON_MZ % 71, 24),
    # 14 -> 7 -> 3 -> 1: the budget computed at sample size 7 (divisor 3.5 * log2(m), not 3) decides how many visits the last three get
    "go5_gmz_s100_m14_b2": ("go", "go5_mz_1bx16", "env_board_size=5:actor_num_simulation=100:zero_num_parallel_games=2:" + GUMBEL % 14 + COMMON_MZ % 72, 24),
    # BASELINE search lengths: configs[1] (400 simulations, the 6b x 256 net) and configs[3] (19x19, 800 simulations): f32 visit counts,
    # chains deeper than 48 levels, nodes with more than 6 visited children
    # Atari MuZero (BASELINE configs[4] search settings: value rescale, reward discount 0.997 — MuZero-paper values) on the
    # synthetic frame source: rewards in the tree, value-bound multiset, #if ATARI init-Q, 601-bin heads, intermediate sequences
    "atari_mz_s20_b2": ("atari", "atari_mz_1bx32", ATARI + "actor_num_simulation=20:zero_num_parallel_games=2:" + COMMON_MZ % 5, 60),
    "atari_mz_s50_b2_det": ("atari", "atari_mz_1bx32", ATARI + "actor_num_simulation=50:zero_num_parallel_games=2:actor_use_dirichlet_noise=false:"
                            "actor_select_action_by_count=true:actor_select_action_by_softmax_count=false:" + COMMON_MZ % 6, 30),
    "atari_mz_s18_gumbel_b2": ("atari", "atari_mz_1bx32", ATARI + "actor_num_simulation=18:zero_num_parallel_games=2:"
                               "actor_use_gumbel=true:actor_use_gumbel_noise=true:actor_gumbel_sample_size=8:actor_gumbel_sigma_visit_c=50:actor_gumbel_sigma_scale_c=0.1:"
                               "actor_use_dirichlet_noise=false:" + COMMON_MZ % 7, 40),
    # intermediate sequences of Atari records (zero_actor_intermediate_sequence_length, 200 by default: atari.h:90): OBS windows, L tags of sent moves
    "atari_mz_seq_s8_b2": ("atari", "atari_mz_1bx32", ATARI + "actor_num_simulation=8:zero_num_parallel_games=2:zero_actor_intermediate_sequence_length=8:"
                           "learner_n_step_return=3:learner_muzero_unrolling_step=2:" + COMMON_MZ % 8, 70),
    "go9_s400_b2": ("go", "go9_az_6bx256", "env_board_size=9:actor_num_simulation=400:zero_num_parallel_games=2:" + COMMON % 81, 6),
    "go19_s800_b2": ("go", "go19_az_1bx16", "env_board_size=19:actor_num_simulation=800:zero_num_parallel_games=2:" + COMMON % 82, 2),
}


def read_case_atari(d, a_size, f_size):
    """muzero_atari recordings: evals carry (6-int header, policy, logits, value, reward); the planes of the root evaluations are
    in roots.bin as integer codes (RGB bytes, action ids); moves carry the value bounds, the children's rewards and what the
    environment answered"""
    ev = np.fromfile(os.path.join(d, "evals.bin"), dtype=np.uint8)
    rec = 24 + 4 * (2 * a_size + 2)
    assert ev.size % rec == 0
    ev = ev.reshape(-1, rec)
    hdr = ev[:, :24].copy().view(np.int32)
    fl = ev[:, 24:].copy().view(np.float32)
    rt = np.fromfile(os.path.join(d, "roots.bin"), dtype=np.uint8).reshape(-1, 8 + f_size)
    rh = rt[:, :8].copy().view(np.int32)
    mv = np.fromfile(os.path.join(d, "moves.bin"), dtype=np.uint8)
    mrec = 4 * 12 + a_size * 32 + 4 * 5
    assert mv.size % mrec == 0
    mv = mv.reshape(-1, mrec)
    mh_i = mv[:, :24].copy().view(np.int32)
    mh_f = mv[:, 24:36].copy().view(np.float32)
    vb_n = mv[:, 36:40].copy().view(np.int32)[:, 0]
    vb = mv[:, 40:48].copy().view(np.float32)
    ch = mv[:, 48:48 + a_size * 32].copy().reshape(-1, a_size, 32)
    tail = mv[:, 48 + a_size * 32:]
    return dict(
        eval_cycle=hdr[:, 0], eval_game=hdr[:, 1], eval_rotation=hdr[:, 2].astype(np.uint8), eval_path_len=hdr[:, 3], eval_leaf_action=hdr[:, 4], eval_path_hash=hdr[:, 5],
        eval_policy=fl[:, :a_size], eval_logits=fl[:, a_size:2 * a_size], eval_value=fl[:, 2 * a_size], eval_reward=fl[:, 2 * a_size + 1],
        root_cycle=rh[:, 0], root_game=rh[:, 1], root_planes=rt[:, 8:].reshape(rt.shape[0], -1),
        move_game=mh_i[:, 0], move_number=mh_i[:, 1], move_action=mh_i[:, 2], move_player=mh_i[:, 3], move_num_children=mh_i[:, 4], move_resign=mh_i[:, 5],
        root_count=mh_f[:, 0], root_mean=mh_f[:, 1], root_value=mh_f[:, 2], bound_size=vb_n, bound_lo=vb[:, 0], bound_hi=vb[:, 1],
        child_action=ch[:, :, 0:4].copy().view(np.int32)[..., 0],
        child_count=ch[:, :, 4:8].copy().view(np.float32)[..., 0], child_mean=ch[:, :, 8:12].copy().view(np.float32)[..., 0],
        child_policy=ch[:, :, 12:16].copy().view(np.float32)[..., 0], child_logit=ch[:, :, 16:20].copy().view(np.float32)[..., 0],
        child_noise=ch[:, :, 20:24].copy().view(np.float32)[..., 0], child_value=ch[:, :, 24:28].copy().view(np.float32)[..., 0],
        child_reward=ch[:, :, 28:32].copy().view(np.float32)[..., 0],
        env_reward=tail[:, 0:4].copy().view(np.float32)[:, 0], env_score=tail[:, 4:8].copy().view(np.float32)[:, 0], env_terminal=tail[:, 8:12].copy().view(np.int32)[:, 0],
        env_seed=tail[:, 12:16].copy().view(np.int32)[:, 0], env_lives=tail[:, 16:20].copy().view(np.int32)[:, 0],
    )


def read_case(d, a_size, f_size, hdr_ints=4):
    ev = np.fromfile(os.path.join(d, "evals.bin"), dtype=np.uint8)
    hb = 4 * hdr_ints
    rec = hb + f_size + 4 * (2 * a_size + 1)
    assert ev.size % rec == 0
    ev = ev.reshape(-1, rec)
    hdr = ev[:, :hb].copy().view(np.int32)
    feats = ev[:, hb:hb + f_size]
    fl = ev[:, hb + f_size:].copy().view(np.float32)
    extra = dict(eval_leaf_action=hdr[:, 4], eval_path_hash=hdr[:, 5]) if hdr_ints == 6 else {}
    mv = np.fromfile(os.path.join(d, "moves.bin"), dtype=np.uint8)
    mrec = 4 * 9 + a_size * 28
    assert mv.size % mrec == 0
    mv = mv.reshape(-1, mrec)
    mh_i = mv[:, :24].copy().view(np.int32)
    mh_f = mv[:, 24:36].copy().view(np.float32)
    ch = mv[:, 36:].copy().reshape(-1, a_size, 28)
    return dict(
        **extra,
        eval_cycle=hdr[:, 0], eval_game=hdr[:, 1], eval_rotation=hdr[:, 2].astype(np.uint8), eval_path_len=hdr[:, 3],
        eval_features=np.packbits(feats, axis=1), eval_policy=fl[:, :a_size], eval_logits=fl[:, a_size:2 * a_size], eval_value=fl[:, 2 * a_size],
        move_game=mh_i[:, 0], move_number=mh_i[:, 1], move_action=mh_i[:, 2], move_player=mh_i[:, 3], move_num_children=mh_i[:, 4], move_resign=mh_i[:, 5],
        root_count=mh_f[:, 0], root_mean=mh_f[:, 1], root_value=mh_f[:, 2],
        child_action=ch[:, :, 0:4].copy().view(np.int32)[..., 0],
        child_count=ch[:, :, 4:8].copy().view(np.float32)[..., 0], child_mean=ch[:, :, 8:12].copy().view(np.float32)[..., 0],
        child_policy=ch[:, :, 12:16].copy().view(np.float32)[..., 0], child_logit=ch[:, :, 16:20].copy().view(np.float32)[..., 0],
        child_noise=ch[:, :, 20:24].copy().view(np.float32)[..., 0], child_value=ch[:, :, 24:28].copy().view(np.float32)[..., 0],
    )


def main(names):
    os.makedirs(OUT, exist_ok=True)
    for name in names:
        binary, net, conf, max_moves = CASES[name]
        with tempfile.TemporaryDirectory() as d:
            conf_full = conf + ":nn_file_name=" + os.path.join(HERE, "_ref", "nets", net + ".pt")
            with open(os.path.join(d, "7x7_seki.db"), "wb") as f:  # KillAllGo's set-up wants its seki table in the working directory: an empty one (see gen_env_golden.py)
                f.write((0).to_bytes(8, "little"))
            res = subprocess.run([os.path.join(HERE, "_ref", "ref_stepper_" + binary), conf_full, d, str(max_moves)], check=True, capture_output=True, text=True, cwd=d)
            meta = dict(line.split() for line in open(os.path.join(d, "meta.txt")))
            if meta["type"] == "muzero_atari":
                data = read_case_atari(d, int(meta["A"]), int(meta["F"]))
            else:
                data = read_case(d, int(meta["A"]), int(meta["F"]), 4 if meta["type"] == "alphazero" else 6)
            data.update(A=int(meta["A"]), F=int(meta["F"]), S=int(meta["S"]), B=int(meta["B"]), conf=conf, net_type=meta["type"], selfplay_lines=np.array(res.stdout.splitlines()))
            np.savez_compressed(os.path.join(OUT, name + ".npz"), **data)
            print(name, "evals", data["eval_game"].size, "moves", data["move_game"].size, "selfplay lines", len(res.stdout.splitlines()),
                  os.path.getsize(os.path.join(OUT, name + ".npz")) // 1024, "KiB")


if __name__ == "__main__":
    main(sys.argv[1:] or list(CASES))
