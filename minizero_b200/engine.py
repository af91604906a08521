"""ctypes binding of libmzb200.so + the host-side mirror of the reference's actor interface for the hot path.

Reference interfaces mirrored (paths relative to /root/reference/minizero):
  network/network.cpp:14-42, create_network.h:11-30  -> Engine.load_network (TorchScript .pt reader)
  network/alphazero_network.h:48-104                 -> Engine.eval_batch
  actor/zero_actor.cpp:51-98                         -> Engine.select / Engine.apply (per-phase parity hooks)
  actor/actor_group.cpp:136-148                      -> Engine.search (whole move, one CUDA graph)
  actor/base_actor.cpp:8-30                          -> Engine.reset_game / Engine.play
"""
import ctypes as C
import os
import subprocess

import numpy as np

GAME_TICTACTOE, GAME_GO, GAME_OTHELLO, GAME_NOGO, GAME_GOMOKU, GAME_HEX, GAME_ATARI, GAME_KILLALLGO = 0, 1, 2, 3, 4, 5, 6, 7
ATARI_FRAME = 3 * 96 * 96  # one screen: RGB bytes, channel-major, 96 x 96 (environment/atari/atari.h:24)
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class EngineError(RuntimeError):
    pass


class _Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("game", C.c_int32), ("board_size", C.c_int32), ("num_games", C.c_int32), ("num_simulation", C.c_int32),
                ("puct_base", C.c_float), ("puct_init", C.c_float), ("reward_discount", C.c_float), ("komi", C.c_float), ("ko_situational", C.c_int32),
                ("dirichlet_epsilon", C.c_float), ("muzero", C.c_int32), ("use_gumbel", C.c_int32), ("gumbel_noise", C.c_int32),
                ("gumbel_sample_size", C.c_int32), ("gumbel_sigma_visit_c", C.c_float), ("gumbel_sigma_scale_c", C.c_float),
                ("gomoku_exactly_five", C.c_int32), ("gomoku_outer_open", C.c_int32), ("hex_swap_rule", C.c_int32), ("value_rescale", C.c_int32),
                ("atari_legal_mask", C.c_uint32), ("think_batch_size", C.c_int32)]


class _NetDims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("num_input_channels", "input_height", "input_width", "num_hidden_channels", "num_blocks", "action_size",
                                        "num_value_hidden_channels", "discrete_value_size", "num_action_feature_channels", "is_muzero")]


class _PlayResult(C.Structure):
    _fields_ = [("applied", C.c_int32), ("terminal", C.c_int32), ("num_legal", C.c_int32), ("turn", C.c_int32), ("eval_score", C.c_float)]


class _RootInfo(C.Structure):
    _fields_ = [("num_children", C.c_int32), ("count", C.c_float), ("mean", C.c_float), ("value", C.c_float)]


_ROOT_INFO_DTYPE = np.dtype([("num_children", np.int32), ("count", np.float32), ("mean", np.float32), ("value", np.float32)])
_PLAY_RESULT_DTYPE = np.dtype([("applied", np.int32), ("terminal", np.int32), ("num_legal", np.int32), ("turn", np.int32), ("eval_score", np.float32)])
assert _ROOT_INFO_DTYPE.itemsize == C.sizeof(_RootInfo) and _PLAY_RESULT_DTYPE.itemsize == C.sizeof(_PlayResult)


def library_path():
    return os.path.join(_ROOT, "minizero_b200", "lib", "libmzb200.so")


NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-ldl"]


def build_library(force=False, verbose=False):
    """nvcc-compile minizero_b200/csrc/engine.cu for sm_100a into minizero_b200/lib/libmzb200.so (in-tree)."""
    src_dir = os.path.join(_ROOT, "minizero_b200", "csrc")
    srcs = [os.path.join(src_dir, f) for f in sorted(os.listdir(src_dir))] + [os.path.join(_ROOT, "include", "mz_b200.h")]
    out = library_path()
    if not force and os.path.exists(out) and os.path.getmtime(out) >= max(os.path.getmtime(s) for s in srcs):
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)
    # MZ_BUILD_EXPERIMENT=1: compile the experiment switches in (tile / stage sweeps, per-phase cycle counters); the release library reads no environment variable
    exp = ["-DMZ_EXPERIMENT"] if os.environ.get("MZ_BUILD_EXPERIMENT") == "1" else []
    cmd = ["nvcc"] + NVCC_FLAGS + exp + (["-Xptxas", "-v"] if verbose else []) + ["-o", out, os.path.join(src_dir, "engine.cu")]
    subprocess.run(cmd, check=True)
    return out


_lib = None


def _load():
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise EngineError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback)")
    lib = C.CDLL(path)
    vp, i32, f32p, u8p, i32p = C.c_void_p, C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.POINTER(C.c_int32)
    lib.mz_last_error.restype = C.c_char_p
    lib.mz_create.argtypes = [C.POINTER(_Config), C.POINTER(vp)]
    lib.mz_destroy.argtypes = [vp]
    lib.mz_destroy.restype = None
    lib.mz_action_size.argtypes = [vp]
    lib.mz_num_features.argtypes = [vp]
    lib.mz_net_configure.argtypes = [vp, C.POINTER(_NetDims)]
    lib.mz_net_set_tensor.argtypes = [vp, C.c_char_p, f32p, C.c_int64]
    lib.mz_net_finalize.argtypes = [vp]
    lib.mz_net_finalize_empty.argtypes = [vp]
    lib.mz_net_blob.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_int64)]
    lib.mz_eval_batch.argtypes = [vp, f32p, i32, f32p, f32p, f32p]
    lib.mz_eval_initial.argtypes = [vp, f32p, i32, f32p, f32p, f32p, f32p]
    lib.mz_eval_recurrent.argtypes = [vp, f32p, i32p, i32, f32p, f32p, f32p, f32p]
    lib.mz_search_leaf.argtypes = [vp, i32p, i32p, i32p]
    lib.mz_eval_rewards.argtypes = [vp, i32, f32p]
    lib.mz_replay_features.argtypes = [vp, i32p, i32, i32p, u8p, i32, f32p]
    lib.mz_atari_observe.argtypes = [vp, i32p, u8p]
    lib.mz_get_root_rewards.argtypes = [vp, f32p, i32p, f32p, f32p]
    lib.mz_search_apply_reward.argtypes = [vp, f32p, f32p, f32p, f32p, f32p]
    lib.mz_gumbel_best_actions.argtypes = [vp, i32p]
    lib.mz_reset_game.argtypes = [vp, i32]
    lib.mz_play.argtypes = [vp, i32p, C.POINTER(_PlayResult)]
    lib.mz_play_max_count.argtypes = [vp, i32, i32p, C.POINTER(_PlayResult)]
    lib.mz_sync.argtypes = [vp]
    lib.mz_timer_begin.argtypes = [vp]
    lib.mz_timer_end.argtypes = [vp, f32p]
    lib.mz_get_roots.argtypes = [vp, C.POINTER(_RootInfo), i32p] + [f32p] * 6
    lib.mz_search_select.argtypes = [vp, u8p, f32p, i32p]
    lib.mz_search_apply.argtypes = [vp, f32p, f32p, f32p, f32p]
    lib.mz_search_set_inputs.argtypes = [vp, u8p, f32p]
    lib.mz_search_run.argtypes = [vp, i32, f32p]
    lib.mz_profile_kernels.argtypes = [vp, i32, f32p, f32p, f32p]
    lib.mz_debug_tree_timing.argtypes = [vp, C.POINTER(C.c_uint64)]
    lib.mz_debug_tower_timing.argtypes = [vp, C.POINTER(C.c_uint64), i32]
    lib.mz_conv_layers_per_launch.argtypes = [vp]
    lib.mz_think_steps.argtypes = [vp]
    lib.mz_tower_is_cooperative.argtypes = [vp]
    lib.mz_tower_is_wide.argtypes = [vp]
    lib.mz_set_tower_cooperative.argtypes = [vp, i32]
    lib.mz_launch_count.argtypes = [vp]
    lib.mz_launch_count.restype = C.c_int64
    _lib = lib
    return lib


EXPORTS = ["mz_create", "mz_destroy", "mz_last_error", "mz_action_size", "mz_num_features", "mz_net_configure", "mz_net_set_tensor", "mz_net_finalize",
           "mz_net_blob", "mz_net_finalize_empty", "mz_eval_batch", "mz_reset_game", "mz_play", "mz_get_roots", "mz_search_select", "mz_search_apply",
           "mz_search_set_inputs", "mz_search_run", "mz_profile_kernels", "mz_launch_count", "mz_play_max_count", "mz_sync", "mz_timer_begin", "mz_timer_end", "mz_debug_tree_timing", "mz_debug_tower_timing", "mz_conv_layers_per_launch",
           "mz_eval_initial", "mz_eval_recurrent", "mz_search_leaf", "mz_gumbel_best_actions", "mz_eval_rewards", "mz_atari_observe", "mz_get_root_rewards",
           "mz_search_apply_reward", "mz_replay_features", "mz_think_steps", "mz_tower_is_cooperative", "mz_set_tower_cooperative", "mz_tower_is_wide"]


def _fp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))


def _u8(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_uint8))


def _i32(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int32))


class Engine:
    """One GPU's share of the self-play games: node pools, environments and the network, all resident in HBM."""

    def __init__(self, game, board_size, num_games, num_simulation, device=0, puct_base=19652.0, puct_init=1.25, reward_discount=1.0, komi=7.5,
                 ko_situational=False, dirichlet_epsilon=0.25, muzero=0, use_gumbel=0, gumbel_noise=0, gumbel_sample_size=16, gumbel_sigma_visit_c=50.0,
                 gumbel_sigma_scale_c=1.0, gomoku_exactly_five=True, gomoku_outer_open=False, hex_swap_rule=True, value_rescale=0,
                 atari_legal_mask=0b1111111101, think_batch_size=0):
        self.lib = _load()
        cfg = _Config(device, game, board_size, num_games, num_simulation, puct_base, puct_init, reward_discount, komi, int(ko_situational), dirichlet_epsilon,
                      int(muzero), int(use_gumbel), int(gumbel_noise), int(gumbel_sample_size), gumbel_sigma_visit_c, gumbel_sigma_scale_c,
                      int(gomoku_exactly_five), int(gomoku_outer_open), int(hex_swap_rule), int(value_rescale), int(atari_legal_mask), int(think_batch_size))
        self.muzero = bool(muzero)
        self.atari = (game == GAME_ATARI)
        self.game, self.board_size = game, (3 if game == GAME_TICTACTOE else (6 if game == GAME_ATARI else board_size))
        h = C.c_void_p()
        self.h = None
        self._check(self.lib.mz_create(C.byref(cfg), C.byref(h)))
        self.h = h
        # console think() (think_batch_size = K > 1): num_games trees, K lanes each; every per-game array of the C ABI has trees * K entries, lane-major
        self.K = int(think_batch_size) if think_batch_size > 1 else 0
        self.trees = num_games
        self.B, self.S = num_games * max(1, self.K), num_simulation
        self.A = self.lib.mz_action_size(self.h)
        self.F = self.lib.mz_num_features(self.h)
        self.terminal = [False] * num_games
        self.last_play = None

    def _check(self, rc):
        if rc != 0:
            raise EngineError(f"libmzb200 error {rc}: {self.lib.mz_last_error().decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.mz_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    # ---- network ------------------------------------------------------------------------------
    def load_network(self, source):
        """source: path of a TorchScript .pt written by the reference's learner (or oracle/gen_nets.py),
        or a (dims dict, state_dict of numpy arrays) pair."""
        if isinstance(source, str):
            import torch  # host-side .pt reader only (the reference uses libtorch for the same, network.cpp:21)
            m = torch.jit.load(source, map_location="cpu")
            dims = dict(num_input_channels=m.get_num_input_channels(), input_height=m.get_input_channel_height(), input_width=m.get_input_channel_width(),
                        num_hidden_channels=m.get_num_hidden_channels(), num_blocks=m.get_num_blocks(), action_size=m.get_action_size(),
                        num_value_hidden_channels=m.get_num_value_hidden_channels(), discrete_value_size=m.get_discrete_value_size())
            if m.get_type_name() in ("muzero", "muzero_atari"):  # network/muzero_network.h:46-52
                dims.update(num_action_feature_channels=m.get_num_action_feature_channels(), is_muzero=1)
            state = {k: v.detach().float().contiguous().numpy() for k, v in m.state_dict().items() if not k.endswith("num_batches_tracked")}
        else:
            dims, state = source
        nd = _NetDims(*[int(dims.get(n, 0)) for n, _ in _NetDims._fields_])
        self._check(self.lib.mz_net_configure(self.h, C.byref(nd)))
        for k, v in state.items():
            a = np.ascontiguousarray(v, np.float32)
            self._check(self.lib.mz_net_set_tensor(self.h, k.encode(), _fp(a), a.size))
        self._check(self.lib.mz_net_finalize(self.h))
        self.net_dims = dims

    def configure_network_empty(self, dims):
        nd = _NetDims(*[int(dims.get(n, 0)) for n, _ in _NetDims._fields_])
        self._check(self.lib.mz_net_configure(self.h, C.byref(nd)))
        self._check(self.lib.mz_net_finalize_empty(self.h))
        self.net_dims = dims

    def weight_blob(self):
        p, n = C.c_void_p(), C.c_int64()
        self._check(self.lib.mz_net_blob(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def eval_batch(self, features):
        f = np.ascontiguousarray(features, np.float32).reshape(-1, self.F)
        n = f.shape[0]
        pol, lg, val = np.zeros((n, self.A), np.float32), np.zeros((n, self.A), np.float32), np.zeros(n, np.float32)
        self._check(self.lib.mz_eval_batch(self.h, _fp(f), n, _fp(pol), _fp(lg), _fp(val)))
        return pol, lg, val

    def eval_initial(self, features):
        """MuZeroNetwork initial inference: policy, logits, value, scaled hidden state [n][Ch*H*W]"""
        f = np.ascontiguousarray(features, np.float32).reshape(-1, self.F)
        n = f.shape[0]
        hsz = int(self.net_dims["num_hidden_channels"]) * (36 if self.atari else int(self.net_dims["input_height"]) * int(self.net_dims["input_width"]))
        pol, lg, val, hid = np.zeros((n, self.A), np.float32), np.zeros((n, self.A), np.float32), np.zeros(n, np.float32), np.zeros((n, hsz), np.float32)
        self._check(self.lib.mz_eval_initial(self.h, _fp(f), n, _fp(pol), _fp(lg), _fp(val), _fp(hid)))
        return pol, lg, val, hid

    def eval_recurrent(self, hidden, actions):
        """MuZeroNetwork recurrent inference on (hidden state, action id) pairs"""
        h = np.ascontiguousarray(hidden, np.float32)
        n = h.shape[0]
        a = np.ascontiguousarray(actions, np.int32)
        pol, lg, val, hid = np.zeros((n, self.A), np.float32), np.zeros((n, self.A), np.float32), np.zeros(n, np.float32), np.zeros_like(h)
        self._check(self.lib.mz_eval_recurrent(self.h, _fp(h), _i32(a), n, _fp(pol), _fp(lg), _fp(val), _fp(hid)))
        return pol, lg, val, hid

    def eval_rewards(self, n):
        """reward head output of the last eval_recurrent (muzero_atari; after the expectation over the bins and invertValue)"""
        r = np.zeros(n, np.float32)
        self._check(self.lib.mz_eval_rewards(self.h, n, _fp(r)))
        return r

    # ---- learner data path --------------------------------------------------------------------------------
    def replay_features(self, actions, positions, rotations=None):
        """BaseEnvLoader::getFeatures for a batch of (record, position, rotation) samples: actions [n][max_len] (-1 padded)"""
        a = np.ascontiguousarray(actions, np.int32)
        n, max_len = a.shape
        pos = np.ascontiguousarray(positions, np.int32)
        rot = None if rotations is None else np.ascontiguousarray(rotations, np.uint8)
        out = np.zeros((n, self.F), np.float32)
        self._check(self.lib.mz_replay_features(self.h, _i32(a), max_len, _i32(pos), _u8(rot), n, _fp(out)))
        return out

    # ---- Atari: the emulator stays with the caller --------------------------------------------------
    def observe_all(self, actions, frames):
        """actions [B]: >= 0 the action the emulator just executed, -1 the first screen after a reset, -2 nothing for this game;
        frames uint8 [B][3][96][96]"""
        a = np.ascontiguousarray(actions, np.int32)
        f = np.ascontiguousarray(frames, np.uint8).reshape(self.B, ATARI_FRAME)
        self._check(self.lib.mz_atari_observe(self.h, _i32(a), _u8(f)))

    def observe(self, g, action, frame, terminal=False):
        a = np.full(self.B, -2, np.int32)
        a[g] = action
        f = np.zeros((self.B, ATARI_FRAME), np.uint8)
        f[g] = np.ascontiguousarray(frame, np.uint8).reshape(-1)
        self.observe_all(a, f)
        self.terminal[g] = bool(terminal)

    # ---- per-phase hooks (per-phase parity hooks) ------------------------
    def select(self, rotations=None, want_features=True):
        rot = None if rotations is None else np.ascontiguousarray(rotations, np.uint8)
        feats = np.zeros((self.B, self.F), np.float32) if want_features else None
        self._path_len = np.zeros(self.B, np.int32)
        self._check(self.lib.mz_search_select(self.h, _u8(rot), _fp(feats), _i32(self._path_len)))
        self._roots = None
        self._leaf = None
        return feats

    def think_select(self, K, rotations=None):
        """selection half of one batched think() step: (planes [K][trees][F], path_len [K][trees]); path_len > 0: evaluate, < 0: duplicate leaf, 0: unused lane"""
        assert K == self.K
        feats = self.select(rotations)
        return feats.reshape(K, self.trees, self.F), self._path_len.reshape(K, self.trees).copy()

    def think_apply(self, policy, logits, value, noise=None):
        nz = None
        if noise is not None:
            nz = np.zeros((self.B, self.A), np.float32)
            nz[:self.trees] = noise
        self.apply(np.asarray(policy).reshape(self.B, self.A), np.asarray(logits).reshape(self.B, self.A), np.asarray(value).reshape(self.B), nz)

    def think_steps(self):
        return int(self.lib.mz_think_steps(self.h))

    def _leaf_info(self):
        if getattr(self, "_leaf", None) is None:
            ps, la, pa = np.zeros(self.B, np.int32), np.zeros(self.B, np.int32), np.zeros((self.B, self.S + 2), np.int32)
            self._check(self.lib.mz_search_leaf(self.h, _i32(ps), _i32(la), _i32(pa)))
            self._leaf = (ps, la, pa)
        return self._leaf

    def leaf_action(self, g):
        return int(self._leaf_info()[1][g])

    def leaf_parent_slot(self, g):
        return int(self._leaf_info()[0][g])

    def path_hash(self, g):
        """FNV-1a over the action ids of the selected path (same formula as oracle/drivers/ref_stepper.cpp)"""
        h = 2166136261
        for a in self._leaf_info()[2][g][1:self.path_len(g)]:
            h = ((h ^ (int(a) & 0xffffffff)) * 16777619) & 0xffffffff
        return h & 0x7fffffff

    def gumbel_best_action(self, g):
        return int(self.gumbel_best_actions()[g])

    def gumbel_best_actions(self):
        out = np.zeros(self.B, np.int32)
        self._check(self.lib.mz_gumbel_best_actions(self.h, _i32(out)))
        return out

    def apply(self, policy, logits, value, noise=None, reward=None):
        p, l, v = (np.ascontiguousarray(x, np.float32) for x in (policy, logits, value))
        nz = None if noise is None else np.ascontiguousarray(noise, np.float32)
        rw = None if reward is None else np.ascontiguousarray(reward, np.float32)
        self._check(self.lib.mz_search_apply_reward(self.h, _fp(p), _fp(l), _fp(v), _fp(rw), _fp(nz)))
        self._roots = None

    def path_len(self, g):
        return int(self._path_len[g])

    def get_roots(self):
        B, A = self.B, self.A
        info = np.zeros(B, _ROOT_INFO_DTYPE)  # mz_root_info[B] without a Python loop per game
        out = dict(action=np.empty((B, A), np.int32))
        for n in ("count", "mean", "policy", "logit", "noise", "value"):
            out[n] = np.empty((B, A), np.float32)
        self._check(self.lib.mz_get_roots(self.h, info.ctypes.data_as(C.POINTER(_RootInfo)), _i32(out["action"]),
                                          *[_fp(out[n]) for n in ("count", "mean", "policy", "logit", "noise", "value")]))
        out["num_children"], out["root_count"], out["root_mean"], out["root_value"] = info["num_children"], info["count"], info["mean"], info["value"]
        out["reward"], out["bound_size"] = np.zeros((B, A), np.float32), np.zeros(B, np.int32)
        out["bound_lo"], out["bound_hi"] = np.zeros(B, np.float32), np.zeros(B, np.float32)
        self._check(self.lib.mz_get_root_rewards(self.h, _fp(out["reward"]), _i32(out["bound_size"]), _fp(out["bound_lo"]), _fp(out["bound_hi"])))
        return out

    def _cached_roots(self):
        if getattr(self, "_roots", None) is None:
            self._roots = self.get_roots()
        return self._roots

    def sims_done(self, g):
        return int(self._cached_roots()["root_count"][g])

    def root(self, g):
        r = self._cached_roots()
        d = dict(num_children=int(r["num_children"][g]), root_count=float(r["root_count"][g]), root_mean=float(r["root_mean"][g]), root_value=float(r["root_value"][g]))
        for n in ("action", "count", "mean", "policy", "logit", "noise", "value", "reward"):
            d[n] = r[n][g]
        d.update(bound_size=int(r["bound_size"][g]), bound_lo=float(r["bound_lo"][g]), bound_hi=float(r["bound_hi"][g]))
        return d

    # ---- games ----------------------------------------------------------------------------------
    def play_all(self, actions):
        a = np.ascontiguousarray(actions, np.int32)
        res = np.zeros(self.B, _PLAY_RESULT_DTYPE)
        self._check(self.lib.mz_play(self.h, _i32(a), res.ctypes.data_as(C.POINTER(_PlayResult))))
        self._roots = None
        out = {k: res[k] for k in ("applied", "terminal", "num_legal", "turn", "eval_score")}
        for g in np.nonzero(a >= 0)[0]:
            self.terminal[g] = bool(out["terminal"][g])
        self.last_play = out
        return out

    def play(self, g, action):
        a = np.full(self.B, -1, np.int32)
        a[g] = action
        return int(self.play_all(a)["applied"][g])

    def root_terminal(self, g):
        return self.terminal[g]

    def reset_game(self, g=-1):
        self._check(self.lib.mz_reset_game(self.h, g))
        self._roots = None
        if g < 0:
            self.terminal = [False] * self.B
        else:
            self.terminal[g] = False

    # ---- whole-move search --------------------------------------------------------------------------
    def set_search_inputs(self, rotations=None, noise=None):
        rot = None if rotations is None else np.ascontiguousarray(rotations, np.uint8).reshape(self.S + 1, self.B)
        nz = None if noise is None else np.ascontiguousarray(noise, np.float32).reshape(self.B, self.A)
        self._keep = (rot, nz)
        self._check(self.lib.mz_search_set_inputs(self.h, _u8(rot), _fp(nz)))

    def search(self, num_evals=0, wait=True):
        """one whole move search for every game; returns the CUDA-event milliseconds when wait is True"""
        ms = C.c_float(0)
        self._check(self.lib.mz_search_run(self.h, num_evals, C.byref(ms) if wait else None))
        self._roots = None
        return ms.value if wait else None

    def play_max_count(self, auto_reset=True, read_back=True):
        self._roots = None
        if not read_back:
            self._check(self.lib.mz_play_max_count(self.h, int(auto_reset), None, None))
            return None
        acts = np.zeros(self.B, np.int32)
        res = (_PlayResult * self.B)()
        self._check(self.lib.mz_play_max_count(self.h, int(auto_reset), _i32(acts), res))
        return dict(action=acts, applied=np.array([r.applied for r in res], np.int32), terminal=np.array([r.terminal for r in res], np.int32),
                    num_legal=np.array([r.num_legal for r in res], np.int32), eval_score=np.array([r.eval_score for r in res], np.float32))

    def sync(self):
        self._check(self.lib.mz_sync(self.h))

    def timer_begin(self):
        self._check(self.lib.mz_timer_begin(self.h))

    def timer_end(self):
        ms = C.c_float(0)
        self._check(self.lib.mz_timer_end(self.h, C.byref(ms)))
        return ms.value

    def profile_kernels(self, iters=20):
        c, t, h = C.c_float(0), C.c_float(0), C.c_float(0)
        self._check(self.lib.mz_profile_kernels(self.h, iters, C.byref(c), C.byref(t), C.byref(h)))
        return dict(conv_ms=c.value, tree_ms=t.value, heads_ms=h.value)

    def tree_timing(self):
        out = np.zeros((self.B, 16), np.uint64)
        self._check(self.lib.mz_debug_tree_timing(self.h, out.ctypes.data_as(C.POINTER(C.c_uint64))))
        return out

    def tower_timing(self):
        out = np.zeros((160, 8), np.uint64)
        n = self.lib.mz_debug_tower_timing(self.h, out.ctypes.data_as(C.POINTER(C.c_uint64)), 160)
        if n < 0:
            self._check(n)
        return out[:n]

    def set_tower_cooperative(self, on):
        """profiler runs switch the cooperative launch off (see include/mz_b200.h); results do not depend on it"""
        self._check(self.lib.mz_set_tower_cooperative(self.h, int(on)))

    def tower_is_wide(self):
        return int(self.lib.mz_tower_is_wide(self.h))

    def tower_is_cooperative(self):
        return int(self.lib.mz_tower_is_cooperative(self.h))

    def conv_layers_per_launch(self):
        return int(self.lib.mz_conv_layers_per_launch(self.h))

    def launch_count(self):
        return int(self.lib.mz_launch_count(self.h))
