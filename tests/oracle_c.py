"""ctypes binding of oracle/liboracle.so (the C restatement of the reference path; test-only)."""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")
_lib = None


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(_DIR, "liboracle.so")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(_DIR, "oracle.c")):
            subprocess.run(["make", "-s", "-C", _DIR], check=True)
        L = C.CDLL(so)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_uint32]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_reset.argtypes = [C.c_void_p]
        L.orc_step.restype = C.c_int
        L.orc_step.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
        L.orc_shaped.restype = C.c_float
        L.orc_shaped.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double]
        L.orc_features.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 4
        L.orc_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_double)]
        L.orc_packages.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_rollout_step.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 7
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class COracleEnv:
    def __init__(self, grid, T, R, P, seed, move_cost=-0.01, delivery_reward=10.0, delay_reward=1.0):
        self.grid = np.ascontiguousarray(grid, dtype=np.uint8)
        self.H, self.W = self.grid.shape
        self.R, self.P, self.T = R, P, T
        self.h = lib().orc_create(_p(self.grid), self.H, self.W, R, P, T, move_cost, delivery_reward, delay_reward, seed)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_destroy(self.h)
            self.h = None

    def reset(self):
        lib().orc_reset(self.h)

    def step(self, actions):
        a = np.ascontiguousarray(actions, dtype=np.uint8)
        r = C.c_double()
        d = lib().orc_step(self.h, _p(a), C.byref(r))
        return r.value, bool(d)

    def shaped(self, actions, consts, reward):
        a = np.ascontiguousarray(actions, dtype=np.uint8)
        k = np.ascontiguousarray(consts, dtype=np.float32)
        return np.float32(lib().orc_shaped(self.h, _p(a), _p(k), reward))

    def state(self):
        rb = np.zeros((self.R, 3), np.int32)
        st = np.zeros(max(self.P, 1), np.uint8)
        t, tot = C.c_int32(), C.c_double()
        lib().orc_state(self.h, _p(rb), _p(st), C.byref(t), C.byref(tot))
        return rb, st[:self.P], t.value, tot.value

    def packages(self):
        pk = np.zeros((max(self.P, 1), 4), np.int32)
        lib().orc_packages(self.h, _p(pk))
        return pk[:self.P]

    def features(self, K_o, K_p, MR, MP):
        HW = self.H * self.W
        obs = np.empty((self.R, 6, self.H, self.W), np.float32)
        vec = np.empty((self.R, 6 + 5 * K_o + 5 * K_p + 1), np.float32)
        g = np.empty((4, self.H, self.W), np.float32)
        gv = np.empty(6 * MR + 7 * MP + 1, np.float32)
        lib().orc_features(self.h, K_o, K_p, MR, MP, _p(obs), _p(vec), _p(g), _p(gv))
        return obs, vec, g, gv


class COracleVec:
    """N environments stepped by one C call per MAPPO rollout iteration."""

    def __init__(self, n, grid, T, R, P, base_seed, K_o, K_p, MR, MP, consts, first_env_index=0):
        self.envs = [COracleEnv(grid, T, R, P, base_seed + first_env_index + i) for i in range(n)]
        self.n, self.R, self.H, self.W = n, R, self.envs[0].H, self.envs[0].W
        self.K = (K_o, K_p, MR, MP)
        self.consts = np.ascontiguousarray(consts, dtype=np.float32)
        self.handles = (C.c_void_p * n)(*[e.h for e in self.envs])
        D, Dg = 6 + 5 * K_o + 5 * K_p + 1, 6 * MR + 7 * MP + 1
        self.reward = np.zeros(n, np.float64)
        self.shaped = np.zeros(n, np.float32)
        self.done = np.zeros(n, np.uint8)
        self.obs = np.zeros((n, R, 6, self.H, self.W), np.float32)
        self.vec = np.zeros((n, R, D), np.float32)
        self.gstate = np.zeros((n, 4, self.H, self.W), np.float32)
        self.gvec = np.zeros((n, Dg), np.float32)

    def reset(self):
        for e in self.envs:
            e.reset()

    def rollout_step(self, actions, features=True):
        a = np.ascontiguousarray(actions, dtype=np.uint8)
        f = (self.obs, self.vec, self.gstate, self.gvec) if features else (None,) * 4
        lib().orc_rollout_step(self.handles, self.n, _p(a), _p(self.consts), *self.K, _p(self.reward), _p(self.shaped),
                               _p(self.done), *[_p(x) for x in f])
        return self.reward, self.shaped, self.done
