"""ctypes binding of the CPU oracle (oracle/port/*.c -> oracle/liboracle.so). TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MAX_ACTIONS = 19 * 19 + 1
GAME_TICTACTOE, GAME_GO, GAME_OTHELLO, GAME_NOGO, GAME_GOMOKU, GAME_HEX, GAME_ATARI, GAME_KILLALLGO = 0, 1, 2, 3, 4, 5, 6, 7
ATARI_LEGAL_MASK = 0b1111111101  # ms_pacman's minimal action set: NOOP, UP, RIGHT, LEFT, DOWN and the four diagonals (ALE ids 0, 2..9)


class Config(C.Structure):
    _fields_ = [("game", C.c_int32), ("board_size", C.c_int32), ("num_games", C.c_int32), ("num_simulation", C.c_int32),
                ("puct_base", C.c_float), ("puct_init", C.c_float), ("reward_discount", C.c_float), ("komi", C.c_float),
                ("ko_situational", C.c_int32), ("value_rescale", C.c_int32), ("dirichlet_epsilon", C.c_float),
                ("muzero", C.c_int32), ("use_gumbel", C.c_int32), ("gumbel_noise", C.c_int32), ("gumbel_sample_size", C.c_int32),
                ("gumbel_sigma_visit_c", C.c_float), ("gumbel_sigma_scale_c", C.c_float), ("gomoku_flags", C.c_int32), ("atari_legal_mask", C.c_uint32)]


class RootOut(C.Structure):
    _fields_ = [("num_children", C.c_int32), ("count", C.c_float), ("mean", C.c_float), ("value", C.c_float),
                ("action", C.c_int32 * MAX_ACTIONS)] + [(n, C.c_float * MAX_ACTIONS) for n in ("c_count", "c_mean", "c_policy", "c_logit", "c_noise", "c_value")]


def default_config(game, board_size, num_games, num_simulation):
    # defaults of config/configuration.cpp:13-28,80
    return Config(game, board_size, num_games, num_simulation, 19652.0, 1.25, 1.0, 7.5, 0, 0, 0.25, 0, 0, 0, 16, 50.0, 1.0, 1 | 4, ATARI_LEGAL_MASK)  # exactly five, hex swap rule


def conf_overrides(conf):
    """search settings of a golden case's conf string -> Config / Engine keyword overrides (config/configuration.cpp:95-195)"""
    kv = dict(item.split("=", 1) for item in str(conf).split(":") if "=" in item)
    true = lambda k, default: kv.get(k, default) == "true"
    extra = {}
    if "actor_mcts_value_rescale" in kv or "actor_mcts_reward_discount" in kv:  # Atari MuZero settings
        extra = dict(value_rescale=int(true("actor_mcts_value_rescale", "false")), reward_discount=float(kv.get("actor_mcts_reward_discount", 1.0)))
    return dict(**extra, muzero=int(kv.get("nn_type_name", "alphazero") == "muzero"), use_gumbel=int(true("actor_use_gumbel", "false")),
                gumbel_noise=int(true("actor_use_gumbel_noise", "false") and not true("actor_use_dirichlet_noise", "true")),
                gumbel_sample_size=int(kv.get("actor_gumbel_sample_size", 16)), gumbel_sigma_visit_c=float(kv.get("actor_gumbel_sigma_visit_c", 50)),
                gumbel_sigma_scale_c=float(kv.get("actor_gumbel_sigma_scale_c", 1)))


def load():
    so = os.path.join(ROOT, "oracle", "liboracle.so")
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"], check=True)
    lib = C.CDLL(so)
    vp, i32, f32p, u8p = C.c_void_p, C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_uint8)
    lib.mzo_create.restype = vp
    lib.mzo_create.argtypes = [C.POINTER(Config)]
    lib.mzo_destroy.argtypes = [vp]
    lib.mzo_reset_game.argtypes = [vp, i32]
    lib.mzo_reset_search.argtypes = [vp, i32]
    lib.mzo_select.argtypes = [vp, u8p, f32p]
    lib.mzo_apply.argtypes = [vp, f32p, f32p, f32p, f32p]
    lib.mzo_num_simulation_done.argtypes = [vp, i32]
    lib.mzo_path_len.argtypes = [vp, i32]
    lib.mzo_root.argtypes = [vp, i32, C.POINTER(RootOut)]
    lib.mzo_root_env.restype = vp
    lib.mzo_root_env.argtypes = [vp, i32]
    lib.mzo_play.argtypes = [vp, i32, i32]
    lib.mzo_select_by_max_count.argtypes = [vp, i32]
    for fn in ("mzo_leaf_action", "mzo_leaf_parent_slot", "mzo_path_hash", "mzo_gumbel_best_action"):
        getattr(lib, fn).argtypes = [vp, i32]
    lib.mzo_gumbel_policy.argtypes = [vp, i32, C.POINTER(C.c_int32), f32p]
    lib.mzo_env_is_terminal.argtypes = [vp]
    lib.mzo_env_is_legal.argtypes = [vp, i32, i32]
    lib.mzo_env_eval_score.restype = C.c_float
    lib.mzo_env_eval_score.argtypes = [vp, i32]
    lib.mzo_apply_mz.argtypes = [vp, f32p, f32p, f32p, f32p, f32p]
    lib.mzo_think_select.argtypes = [vp, i32, u8p, f32p, C.POINTER(C.c_int32)]
    lib.mzo_think_apply.argtypes = [vp, f32p, f32p, f32p, f32p]
    lib.mzo_think_leaf.argtypes = [vp, i32, i32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    lib.mzo_root_extra.argtypes = [vp, i32, f32p, C.POINTER(C.c_int32), f32p, f32p]
    lib.mzo_root_normalized_mean.restype = C.c_float
    lib.mzo_root_normalized_mean.argtypes = [vp, i32, i32]
    lib.mzo_atari_observe.argtypes = [vp, i32, i32, u8p, i32]
    lib.mzo_net_create.restype = vp
    lib.mzo_net_create.argtypes = [i32] * 7
    lib.mzo_net_destroy.argtypes = [vp]
    lib.mzo_net_set.argtypes = [vp, C.c_char_p, f32p, C.c_int64]
    lib.mzo_net_forward.argtypes = [vp, f32p, i32, f32p, f32p, f32p]
    return lib


def fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def u8ptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


class OracleSearch:
    """Thin numpy front-end over mzo_batch."""

    def __init__(self, lib, game, board_size, num_games, num_simulation, **overrides):
        self.lib = lib
        self.cfg = default_config(game, board_size, num_games, num_simulation)
        for k, v in overrides.items():
            setattr(self.cfg, k, v)
        self.h = lib.mzo_create(C.byref(self.cfg))
        n = 3 if game == GAME_TICTACTOE else board_size
        self.A = 9 if game == GAME_TICTACTOE else (n * n if game in (GAME_GOMOKU, GAME_HEX) else n * n + 1)
        self.F = (18 if game in (GAME_GO, GAME_NOGO, GAME_KILLALLGO) else 4) * n * n
        self.atari = (game == GAME_ATARI)
        if self.atari:
            self.A, self.F = 18, 32 * 96 * 96
            self._terminal = [False] * num_games
        self.B, self.S = num_games, num_simulation

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.mzo_destroy(self.h)
            self.h = None

    def select(self, rotations=None):
        feats = np.zeros((self.B, self.F), np.float32)
        rot = None if rotations is None else np.ascontiguousarray(rotations, np.uint8)
        self.lib.mzo_select(self.h, None if rot is None else u8ptr(rot), fptr(feats))
        return feats

    def apply(self, policy, logits, value, noise=None, reward=None):
        p, l, v = (np.ascontiguousarray(x, np.float32) for x in (policy, logits, value))
        nz = None if noise is None else np.ascontiguousarray(noise, np.float32)
        rw = None if reward is None else np.ascontiguousarray(reward, np.float32)
        self.lib.mzo_apply_mz(self.h, fptr(p), fptr(l), fptr(v), None if rw is None else fptr(rw), None if nz is None else fptr(nz))

    def think_select(self, K, rotations=None):
        """one batched think() step, selection half: lane-major (features [K][B][F], path_len [K][B])"""
        feats = np.zeros((K, self.B, self.F), np.float32)
        plen = np.zeros((K, self.B), np.int32)
        rot = None if rotations is None else np.ascontiguousarray(rotations, np.uint8)
        self.lib.mzo_think_select(self.h, K, None if rot is None else u8ptr(rot), fptr(feats), plen.ctypes.data_as(C.POINTER(C.c_int32)))
        return feats, plen

    def think_leaf(self, k, g):
        """MuZero: (evaluation slot of the leaf's parent, leaf action) of lane k of tree g; (-1, -1) for the root"""
        ps, a = C.c_int32(0), C.c_int32(0)
        self.lib.mzo_think_leaf(self.h, k, g, C.byref(ps), C.byref(a))
        return ps.value, a.value

    def think_apply(self, policy, logits, value, noise=None):
        p, l, v = (np.ascontiguousarray(x, np.float32) for x in (policy, logits, value))
        nz = None if noise is None else np.ascontiguousarray(noise, np.float32)
        self.lib.mzo_think_apply(self.h, fptr(p), fptr(l), fptr(v), None if nz is None else fptr(nz))

    def observe(self, g, action, frame, terminal=False):
        """Atari: the emulator's answer to `action` (action < 0: the initial screen after a reset), frame = uint8 [3][96][96]"""
        f = np.ascontiguousarray(frame, np.uint8).reshape(-1)
        assert f.size == 3 * 96 * 96
        self.lib.mzo_atari_observe(self.h, g, int(action), u8ptr(f), int(terminal))
        self._terminal[g] = bool(terminal)

    def normalized_mean(self, g, child=-1):
        return float(self.lib.mzo_root_normalized_mean(self.h, g, child))

    def sims_done(self, g):
        return self.lib.mzo_num_simulation_done(self.h, g)

    def path_len(self, g):
        return self.lib.mzo_path_len(self.h, g)

    def leaf_action(self, g):
        return self.lib.mzo_leaf_action(self.h, g)

    def leaf_parent_slot(self, g):
        return self.lib.mzo_leaf_parent_slot(self.h, g)

    def path_hash(self, g):
        return self.lib.mzo_path_hash(self.h, g)

    def gumbel_best_action(self, g):
        return self.lib.mzo_gumbel_best_action(self.h, g)

    def gumbel_policy(self, g):
        a, p = np.zeros(self.A, np.int32), np.zeros(self.A, np.float32)
        n = self.lib.mzo_gumbel_policy(self.h, g, a.ctypes.data_as(C.POINTER(C.c_int32)), fptr(p))
        return a[:n], p[:n]

    def root(self, g):
        out = RootOut()
        self.lib.mzo_root(self.h, g, C.byref(out))
        k = out.num_children
        d = dict(num_children=k, root_count=out.count, root_mean=out.mean, root_value=out.value, action=np.array(out.action[:self.A], np.int32))
        for n in ("c_count", "c_mean", "c_policy", "c_logit", "c_noise", "c_value"):
            d[n[2:]] = np.array(getattr(out, n)[:self.A], np.float32)
        rw, bn, lo, hi = np.zeros(self.A, np.float32), C.c_int32(0), C.c_float(0), C.c_float(0)
        self.lib.mzo_root_extra(self.h, g, fptr(rw), C.byref(bn), C.byref(lo), C.byref(hi))
        d.update(reward=rw, bound_size=bn.value, bound_lo=lo.value, bound_hi=hi.value)
        return d

    def play(self, g, action):
        return self.lib.mzo_play(self.h, g, action)

    def root_terminal(self, g):
        if self.atari:
            return self._terminal[g]
        return bool(self.lib.mzo_env_is_terminal(self.lib.mzo_root_env(self.h, g)))

    def reset_game(self, g):
        self.lib.mzo_reset_game(self.h, g)


class OracleNet:
    """fp32 network restatement (oracle/port/mzo_net.c) with a reference-named state_dict."""

    def __init__(self, lib, dims, state):
        self.lib, self.A = lib, dims["action_size"]
        self.h = lib.mzo_net_create(dims["num_input_channels"], dims["input_height"], dims["input_width"], dims["num_hidden_channels"], dims["num_blocks"],
                                    dims["action_size"], dims["num_value_hidden_channels"])
        for k, v in state.items():
            a = np.ascontiguousarray(v, np.float32)
            assert lib.mzo_net_set(self.h, k.encode(), fptr(a), a.size) == 0

    def forward(self, feats):
        f = np.ascontiguousarray(feats, np.float32)
        n = f.shape[0]
        pol, lg, val = np.zeros((n, self.A), np.float32), np.zeros((n, self.A), np.float32), np.zeros(n, np.float32)
        self.lib.mzo_net_forward(self.h, fptr(f), n, fptr(pol), fptr(lg), fptr(val))
        return pol, lg, val
