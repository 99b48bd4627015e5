"""TEST INFRASTRUCTURE: loader for the plain-C++ build of the per-env device logic
(tests/host_emul/emul.cpp).  Lets the CPU test tier check the *generated kernel source*
(model RHS, symbolic LU, the Rosenbrock stepper, events, outputs) against the oracle without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np

import gym_pvder_b200 as G
from gym_pvder_b200 import _cabi

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "host_emul", "emul.cpp")
_LIB = os.path.join(_HERE, "host_emul", "libpvder_emul.so")
_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    deps = [_SRC] + [os.path.join(_cabi.CSRC, f) for f in os.listdir(_cabi.CSRC) if f.endswith(".cuh")]
    if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < max(os.path.getmtime(d) for d in deps):
        # build to a private name and rename: several processes (the 2-rank gloo test, pytest-xdist) may find the library
        # stale at the same time, and none of them may ever load a half-written file
        tmp = f"{_LIB}.{os.getpid()}.tmp"
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                        "-o", tmp, _SRC], check=True)
        os.replace(tmp, _LIB)
    _lib = C.CDLL(_LIB)
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class EmulVecEnv:
    """Same SoA state layout and step semantics as PVDERVecEnv, run by the C++ emulation."""

    def __init__(self, num_envs, env_offset=0, **kw):
        self.lib = load()
        self.cfg = G.EnvConfig(**kw)
        self.n = num_envs
        self.off = env_offset
        self.ns = self.cfg.n_state
        self.ld = num_envs
        self.sd = np.zeros((_cabi.sd_fields(self.ns), self.ld))
        self.si = np.zeros((_cabi.SI_FIELDS, self.ld), dtype=np.int32)
        self.obs64 = np.zeros((num_envs, 11))
        self.reward = np.zeros(num_envs)
        self.reward_i = np.zeros(num_envs, dtype=np.int32)
        self.done = np.zeros(num_envs, dtype=np.uint8)
        self.vtab = self.stab = None
        self._init = False

    def set_event_tables(self, v, s):
        self.vtab = np.ascontiguousarray(v, dtype=np.float64)
        self.stab = np.ascontiguousarray(s, dtype=np.float64)

    def reset(self):
        self.lib.emul_reset(C.byref(self.cfg.c), _p(self.sd), _p(self.si), C.c_int64(self.ld), 0 if self._init else 1,
                            _p(self.obs64), C.c_int64(self.n), C.c_int64(self.off))
        self._init = True
        return self.obs64.copy()

    def step(self, actions, record=False):
        a = np.ascontiguousarray(actions, dtype=np.int32)
        if record:
            self.traj = np.zeros((self.cfg.c.n_sub_per_step, self.ns + 2, self.n))
            self.lib.emul_set_traj(_p(self.traj))
        else:
            self.lib.emul_set_traj(None)
        self.lib.emul_step(C.byref(self.cfg.c), _p(self.sd), _p(self.si), C.c_int64(self.ld), _p(a), _p(self.vtab),
                           _p(self.stab), _p(self.obs64), _p(self.reward), _p(self.reward_i), _p(self.done),
                           C.c_int64(self.n), C.c_int64(self.off))
        rew = self.reward_i.copy() if self.cfg.DISCRETE_REWARD else self.reward.copy()
        return self.obs64.copy(), rew, self.done.astype(bool), {}

    def events(self):
        k = max(1, self.cfg.c.ev_count)
        v = np.ones((k, self.ld))
        s = np.full((k, self.ld), 100.0)
        ep = np.ascontiguousarray(self.si[_cabi.SI_EPISODE])
        self.lib.emul_events(C.byref(self.cfg.c), _p(ep), _p(v), _p(s), C.c_int64(self.ld), C.c_int64(self.n),
                             C.c_int64(self.off))
        return v, s


def scheme(cfg):
    """Integrator tableau compiled into the kernels: dict(scheme, gamma, a, c, m, d4, h_gamma)."""
    out = np.zeros(34)
    load().emul_scheme(C.byref(cfg.c), _p(out))
    return dict(scheme=int(out[0]), gamma=out[1], a=out[2:12], c=out[12:27], m=out[27:31], d4=out[31:33], h_gamma=out[33])


def rhs(cfg, y, inp4, frz=0):
    f = np.zeros(cfg.n_state)
    y = np.ascontiguousarray(y, dtype=np.float64)
    i4 = np.ascontiguousarray(inp4, dtype=np.float64)
    load().emul_rhs(C.byref(cfg.c), _p(y), _p(i4), C.c_uint(frz), _p(f))
    return f


def wsolve(cfg, y, inp4, ghinv, b, frz=0):
    b = np.array(b, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    i4 = np.ascontiguousarray(inp4, dtype=np.float64)
    load().emul_wsolve(C.byref(cfg.c), _p(y), _p(i4), C.c_uint(frz), C.c_double(ghinv), _p(b))
    return b


def freeze_bits(cfg, y, inp4):
    y = np.ascontiguousarray(y, dtype=np.float64)
    i4 = np.ascontiguousarray(inp4, dtype=np.float64)
    load().emul_freeze_bits.restype = C.c_uint
    return int(load().emul_freeze_bits(C.byref(cfg.c), _p(y), _p(i4)))


def split_free_path(on):
    """frz == 0 calls use the FREE instantiation (gains from the constant table) when on, else the general one."""
    load().emul_split_free_path(C.c_int(1 if on else 0))


def split_rhs(cfg, y, inp4, frz=0):
    """Right-hand side of the lane-split three-phase model (pvder_split3.cuh), lanes emulated on the host."""
    f = np.zeros(23)
    y = np.ascontiguousarray(y, dtype=np.float64)
    i4 = np.ascontiguousarray(inp4, dtype=np.float64)
    load().emul_split_rhs(C.byref(cfg.c), _p(y), _p(i4), C.c_uint(frz), _p(f))
    return f


def split_wsolve(cfg, y, inp4, ghinv, b, frz=0):
    b = np.array(b, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    i4 = np.ascontiguousarray(inp4, dtype=np.float64)
    load().emul_split_wsolve(C.byref(cfg.c), _p(y), _p(i4), C.c_uint(frz), C.c_double(ghinv), _p(b))
    return b


def split_freeze_bits(cfg, y, inp4):
    y = np.ascontiguousarray(y, dtype=np.float64)
    i4 = np.ascontiguousarray(inp4, dtype=np.float64)
    load().emul_split_freeze_bits.restype = C.c_uint
    return int(load().emul_split_freeze_bits(C.byref(cfg.c), _p(y), _p(i4)))
