"""ctypes front-end of tests/hostsim (search_core.cuh compiled as plain C++, MZ_W == 1). TEST ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def load():
    src = os.path.join(HERE, "hostsim", "hostsim.cpp")
    so = os.path.join(HERE, "hostsim", "libhostsim.so")
    core = os.path.join(HERE, "..", "minizero_b200", "csrc", "search_core.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(core)):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", so, src], check=True)
    lib = C.CDLL(so)
    vp, i32, f32, f32p, u8p, i32p = C.c_void_p, C.c_int32, C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.POINTER(C.c_int32)
    lib.hs_create.restype = vp
    lib.hs_create.argtypes = [i32, i32, i32, i32, f32, f32, f32, f32, f32]
    lib.hs_select.argtypes = [vp, u8p, f32p]
    lib.hs_apply.argtypes = [vp, f32p, f32p, f32p, f32p]
    lib.hs_path_len.argtypes = [vp, i32]
    lib.hs_sims_done.argtypes = [vp, i32]
    lib.hs_root.argtypes = [vp, i32, i32p, f32p]
    lib.hs_play.argtypes = [vp, i32, i32, i32p, f32p]
    lib.hs_reset_game.argtypes = [vp, i32]
    lib.hs_destroy.argtypes = [vp]
    lib.hs_check_level_variants.argtypes = [vp, i32]
    lib.hs_sort_matches_std.argtypes = [i32, f32p, i32p]
    lib.hs_set_options.argtypes = [i32, i32, i32, i32, f32, f32]
    lib.hs_set_atari_options.argtypes = [i32, i32]
    lib.hs_apply_mz.argtypes = [vp, f32p, f32p, f32p, f32p, f32p]
    lib.hs_set_think.argtypes = [i32]
    lib.hs_think_select.argtypes = [vp, u8p, f32p, i32p]
    lib.hs_think_apply.argtypes = [vp, f32p, f32p, f32p, f32p]
    lib.hs_atari_observe.argtypes = [vp, i32, i32]
    lib.hs_root_extra.argtypes = [vp, i32, f32p, i32p]
    for fn in ("hs_leaf_action", "hs_leaf_parent_slot", "hs_path_hash", "hs_gumbel_best_action"):
        getattr(lib, fn).argtypes = [vp, i32]
    return lib


def root_dict(A, out_i, out_f):
    k = int(out_i[0])
    d = dict(num_children=k, root_count=float(out_f[0]), root_mean=float(out_f[1]), root_value=float(out_f[2]), action=out_i[1:1 + A].copy())
    for j, n in enumerate(("count", "mean", "policy", "logit", "noise", "value")):
        d[n] = out_f[3 + j * A:3 + (j + 1) * A].copy()
    return d


class HostSimSearch:
    def __init__(self, lib, game, board_size, num_games, num_simulation, muzero=0, use_gumbel=0, gumbel_noise=0, gumbel_sample_size=16,
                 gumbel_sigma_visit_c=50.0, gumbel_sigma_scale_c=1.0, value_rescale=0, reward_discount=1.0, atari_legal_mask=0b1111111101, think_k=0):
        self.lib = lib
        n = 3 if game == 0 else board_size
        self.A = 9 if game == 0 else (n * n if game in (4, 5) else n * n + 1)
        self.F = (18 if game in (1, 3, 7) else 4) * n * n
        if game == 6:  # Atari MuZero: 18 actions, planes from the device's screen ring (not produced by the host build)
            self.A, self.F = 18, 0
        self.B, self.S = num_games, num_simulation
        lib.hs_set_options(muzero, use_gumbel, gumbel_noise, gumbel_sample_size, gumbel_sigma_visit_c, gumbel_sigma_scale_c)
        lib.hs_set_atari_options(value_rescale, atari_legal_mask)
        lib.hs_set_think(think_k)
        self.K = think_k
        self.h = lib.hs_create(game, n, num_games, num_simulation, 19652.0, 1.25, reward_discount, 7.5, 0.25)
        self.terminal = [False] * num_games

    def select(self, rotations=None):
        feats = np.zeros((self.B, self.F), np.float32)
        rot = None if rotations is None else np.ascontiguousarray(rotations, np.uint8)
        self.lib.hs_select(self.h, None if rot is None else rot.ctypes.data_as(C.POINTER(C.c_uint8)), feats.ctypes.data_as(C.POINTER(C.c_float)))
        return feats

    def apply(self, policy, logits, value, noise=None, reward=None):
        fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
        p, l, v = (np.ascontiguousarray(x, np.float32) for x in (policy, logits, value))
        nz = None if noise is None else np.ascontiguousarray(noise, np.float32)
        rw = None if reward is None else np.ascontiguousarray(reward, np.float32)
        self.lib.hs_apply_mz(self.h, fp(p), fp(l), fp(v), None if rw is None else fp(rw), None if nz is None else fp(nz))

    def think_select(self, K, rotations=None):
        assert K == self.K
        feats = np.zeros((K, self.B, self.F), np.float32)
        plen = np.zeros((K, self.B), np.int32)
        rot = None if rotations is None else np.ascontiguousarray(rotations, np.uint8)
        self.lib.hs_think_select(self.h, None if rot is None else rot.ctypes.data_as(C.POINTER(C.c_uint8)), feats.ctypes.data_as(C.POINTER(C.c_float)),
                                 plen.ctypes.data_as(C.POINTER(C.c_int32)))
        return feats, plen

    def think_apply(self, policy, logits, value, noise=None):
        fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
        p, l, v = (np.ascontiguousarray(x, np.float32) for x in (policy, logits, value))
        nz = None if noise is None else np.ascontiguousarray(noise, np.float32)
        self.lib.hs_think_apply(self.h, fp(p), fp(l), fp(v), None if nz is None else fp(nz))

    def observe(self, g, action, frame, terminal=False):
        self.lib.hs_atari_observe(self.h, g, int(action))
        self.terminal[g] = bool(terminal)

    def sims_done(self, g):
        return self.lib.hs_sims_done(self.h, g)

    def path_len(self, g):
        return self.lib.hs_path_len(self.h, g)

    def leaf_action(self, g):
        return self.lib.hs_leaf_action(self.h, g)

    def leaf_parent_slot(self, g):
        return self.lib.hs_leaf_parent_slot(self.h, g)

    def path_hash(self, g):
        return self.lib.hs_path_hash(self.h, g)

    def gumbel_best_action(self, g):
        return self.lib.hs_gumbel_best_action(self.h, g)

    def check_level_variants(self, g):
        return self.lib.hs_check_level_variants(self.h, g)

    def root(self, g):
        out_i = np.zeros(1 + self.A, np.int32)
        out_f = np.zeros(3 + 6 * self.A, np.float32)
        self.lib.hs_root(self.h, g, out_i.ctypes.data_as(C.POINTER(C.c_int32)), out_f.ctypes.data_as(C.POINTER(C.c_float)))
        d = root_dict(self.A, out_i, out_f)
        rw, bound = np.zeros(self.A, np.float32), np.zeros(3, np.int32)
        self.lib.hs_root_extra(self.h, g, rw.ctypes.data_as(C.POINTER(C.c_float)), bound.ctypes.data_as(C.POINTER(C.c_int32)))
        d.update(reward=rw, bound_size=int(bound[0]), bound_lo=float(bound[1:2].view(np.float32)[0]), bound_hi=float(bound[2:3].view(np.float32)[0]))
        return d

    def play(self, g, action):
        nl, sc = C.c_int32(0), C.c_float(0)
        r = self.lib.hs_play(self.h, g, action, C.byref(nl), C.byref(sc))
        self.terminal[g] = bool(r & 2)
        self.last_score = sc.value
        return r & 1

    def root_terminal(self, g):
        return self.terminal[g]

    def reset_game(self, g):
        self.lib.hs_reset_game(self.h, g)
        self.terminal[g] = False
