"""Golden vectors of the console search (ZeroActor::think with actor_mcts_think_batch_size = K > 1, zero_actor.cpp:36-49,129-157), recorded from the
UNMODIFIED reference (oracle/_ref/ref_think_*, driver oracle/drivers/ref_think.cpp). Build container only. Output: tests/golden/<case>.npz with, per
batched step, every selection (lane, rotation, path length, the leaf's virtual loss before the selection, bit-packed planes) and every network
output the actor consumed; per search the root child table it ended with. TEST INFRASTRUCTURE ONLY."""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden")
COMMON = "zero_num_threads=1:program_seed=%d:program_auto_seed=false:program_quiet=true:nn_type_name=alphazero:actor_select_action_by_count=true:actor_select_action_by_softmax_count=false"
CASES = {
    # name: (binary, net, conf, moves)
    "think_ttt_s50_k4": ("tictactoe", "ttt_az_2bx32", "actor_num_simulation=50:actor_mcts_think_batch_size=4:" + COMMON % 91, 9),
    "think_go5_s60_k8": ("go", "go5_az_1bx16", "env_board_size=5:actor_num_simulation=60:actor_mcts_think_batch_size=8:" + COMMON % 92, 12),
    # no noise, no rotation: every step of the search is a function of the network alone
    "think_go9_s100_k16_det": ("go", "go9_az_1bx16", "env_board_size=9:actor_num_simulation=100:actor_mcts_think_batch_size=16:actor_use_dirichlet_noise=false:"
                               "actor_use_random_rotation_features=false:" + COMMON % 93, 4),
    # a batch that does not divide the simulations: the last step is short (batch_size = min(K, simulations left))
    "think_go5_s23_k5": ("go", "go5_az_1bx16", "env_board_size=5:actor_num_simulation=23:actor_mcts_think_batch_size=5:" + COMMON % 94, 30),
    # MuZero: the root's initial inference is a batch of one, the recurrent steps select K leaves (zero_actor.cpp:134-135); hidden-state slots in evaluation order
    "think_othello_mz_s30_k6": ("othello", "othello_mz_1bx32", "actor_num_simulation=30:actor_mcts_think_batch_size=6:" + (COMMON % 96).replace("nn_type_name=alphazero", "nn_type_name=muzero"), 14),
    "think_go5_mz_s20_k4": ("go", "go5_mz_1bx16", "env_board_size=5:actor_num_simulation=20:actor_mcts_think_batch_size=4:" + (COMMON % 97).replace("nn_type_name=alphazero", "nn_type_name=muzero"), 16),
}


def read_case(d, A, F):
    ev = open(os.path.join(d, "events.bin"), "rb").read()
    pos = 0
    sel, out, step_of_search = [], [], []
    search, step = 0, -1
    while pos < len(ev):
        kind = int(np.frombuffer(ev, np.int32, 1, pos)[0])
        pos += 4
        if kind == 0:
            bid, rot, plen = (int(x) for x in np.frombuffer(ev, np.int32, 3, pos))
            vl = float(np.frombuffer(ev, np.float32, 1, pos + 12)[0])
            feats = np.frombuffer(ev, np.uint8, F, pos + 16)
            pos += 16 + F
            if bid == 0:
                step += 1
            sel.append((search, step, bid, rot, plen, vl, np.packbits(feats)))
        elif kind == 1:
            bid = int(np.frombuffer(ev, np.int32, 1, pos)[0])
            fl = np.frombuffer(ev, np.float32, 2 * A + 1, pos + 4)
            pos += 4 + 4 * (2 * A + 1)
            out.append((search, step, bid, fl[:A].copy(), fl[A:2 * A].copy(), float(fl[2 * A])))
        else:
            search += 1
    mv = np.fromfile(os.path.join(d, "moves.bin"), dtype=np.uint8)
    mrec = 24 + A * 32
    assert mv.size % mrec == 0
    mv = mv.reshape(-1, mrec)
    mh_i = mv[:, :12].copy().view(np.int32)
    mh_f = mv[:, 12:24].copy().view(np.float32)
    ch = mv[:, 24:].copy().reshape(-1, A, 32)
    col = lambda k, t: ch[:, :, 4 * k:4 * k + 4].copy().view(t)[..., 0]
    return dict(
        sel_search=np.array([s[0] for s in sel]), sel_step=np.array([s[1] for s in sel]), sel_lane=np.array([s[2] for s in sel]),
        sel_rotation=np.array([s[3] for s in sel], np.uint8), sel_path_len=np.array([s[4] for s in sel]), sel_leaf_vloss=np.array([s[5] for s in sel], np.float32),
        sel_features=np.stack([s[6] for s in sel]),
        out_search=np.array([o[0] for o in out]), out_step=np.array([o[1] for o in out]), out_lane=np.array([o[2] for o in out]),
        out_policy=np.stack([o[3] for o in out]), out_logits=np.stack([o[4] for o in out]), out_value=np.array([o[5] for o in out], np.float32),
        move_action=mh_i[:, 0], move_player=mh_i[:, 1], move_num_children=mh_i[:, 2], root_count=mh_f[:, 0], root_mean=mh_f[:, 1], root_value=mh_f[:, 2],
        child_action=col(0, np.int32), child_count=col(1, np.float32), child_mean=col(2, np.float32), child_policy=col(3, np.float32), child_logit=col(4, np.float32),
        child_noise=col(5, np.float32), child_value=col(6, np.float32), child_vloss=col(7, np.float32),
    )


def main(names):
    for name in names:
        binary, net, conf, moves = CASES[name]
        with tempfile.TemporaryDirectory() as d:
            conf_full = conf + ":nn_file_name=" + os.path.join(HERE, "_ref", "nets", net + ".pt")
            subprocess.run([os.path.join(HERE, "_ref", "ref_think_" + binary), conf_full, d, str(moves)], check=True, capture_output=True, text=True)
            meta = dict(line.split() for line in open(os.path.join(d, "meta.txt")))
            data = read_case(d, int(meta["A"]), int(meta["F"]))
            data.update(A=int(meta["A"]), F=int(meta["F"]), S=int(meta["S"]), K=int(meta["K"]), conf=conf)
            np.savez_compressed(os.path.join(OUT, name + ".npz"), **data)
            print(name, "selections", data["sel_lane"].size, "evaluations", data["out_lane"].size, "searches", data["move_action"].size,
                  "duplicates", int((data["sel_leaf_vloss"] != 0).sum()), os.path.getsize(os.path.join(OUT, name + ".npz")) // 1024, "KiB")


if __name__ == "__main__":
    main(sys.argv[1:] or list(CASES))
