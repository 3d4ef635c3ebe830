"""ctypes face of the CPU ORACLE (oracle/dq_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(deepq_decoding_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libdq_oracle.so")
    src = os.path.join(_HERE, "dq_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libdq_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.dqo_create.restype = C.c_void_p
        L.dqo_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                 C.c_int64, C.c_uint64, C.c_int64]
        L.dqo_destroy.argtypes = [C.c_void_p]
        L.dqo_set_noise.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.dqo_set_max_attempts.argtypes = [C.c_void_p, C.c_int]
        L.dqo_set_referee.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.dqo_info.argtypes = [C.c_void_p, C.c_int]
        L.dqo_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.dqo_step.argtypes = [C.c_void_p] + [C.c_void_p] * 6 + [C.c_int]
        L.dqo_random_legal_actions.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        L.dqo_get_env.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 5
        L.dqo_set_hidden.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        L.dqo_syndrome_of.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.dqo_stab_order.argtypes = [C.c_void_p] * 4
        L.dqo_philox.argtypes = [C.c_uint32] * 6 + [C.c_void_p]
        L.dqo_num_threads.restype = C.c_int
        L.dqo_set_num_threads.argtypes = [C.c_int]
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


MODEL = {"X": 0, "DP": 1}


class OracleVecEnv:
    """N independent reference-semantics environments stepped on the CPU."""

    def __init__(self, d, error_model, use_Y, volume_depth, p_phys, p_meas, n_envs, seed, env_id_base=0):
        self.L = lib()
        self.h = self.L.dqo_create(d, MODEL[error_model], int(use_Y), volume_depth, float(p_phys),
                                   float(p_meas), n_envs, seed, env_id_base)
        if not self.h:
            raise ValueError("bad oracle parameters")
        self.d, self.n = d, n_envs
        self.A = self.L.dqo_info(self.h, 0)
        self.Cn = self.L.dqo_info(self.h, 1)
        self.H = self.L.dqo_info(self.h, 2)
        self.ns = self.L.dqo_info(self.h, 3)
        self.B = self.L.dqo_info(self.h, 4)
        self.n3 = self.L.dqo_info(self.h, 5)
        self.n1 = self.L.dqo_info(self.h, 6)
        self.W = self.L.dqo_info(self.h, 7)
        self._luts = None

    def __del__(self):
        if getattr(self, "h", None):
            self.L.dqo_destroy(self.h)
            self.h = None

    def set_max_attempts(self, n):
        """0 = redraw an all-trivial volume for ever (the reference); n > 0 = accept it after n attempts (the product's deviation)."""
        self.L.dqo_set_max_attempts(self.h, int(n))

    def set_noise(self, p_phys, p_meas):
        self.L.dqo_set_noise(self.h, float(p_phys), float(p_meas))

    def set_referee(self, mode, lut_a, lut_b=None):
        lut_a = np.ascontiguousarray(lut_a, dtype=np.uint8)
        lut_b = None if lut_b is None else np.ascontiguousarray(lut_b, dtype=np.uint8)
        self._luts = (lut_a, lut_b)
        self.L.dqo_set_referee(self.h, mode, _p(lut_a), _p(lut_b))

    def reset(self):
        obs = np.empty((self.n, self.Cn, self.H, self.H), np.uint8)
        legal = np.empty((self.n, self.W), np.uint64)
        self.L.dqo_reset(self.h, _p(obs), _p(legal))
        return obs, legal

    def step(self, actions, auto_reset=True, want_obs=True):
        actions = np.ascontiguousarray(actions, dtype=np.int32)
        obs = np.empty((self.n, self.Cn, self.H, self.H), np.uint8) if want_obs else None
        reward = np.empty(self.n, np.float32)
        done = np.empty(self.n, np.uint8)
        life = np.empty(self.n, np.int32)
        legal = np.empty((self.n, self.W), np.uint64)
        self.L.dqo_step(self.h, _p(actions), _p(obs), _p(reward), _p(done), _p(life), _p(legal), int(auto_reset))
        return obs, reward, done, life, legal

    def random_legal_actions(self, legal, step):
        legal = np.ascontiguousarray(legal, dtype=np.uint64)
        act = np.empty(self.n, np.int32)
        self.L.dqo_random_legal_actions(self.h, _p(legal), step, _p(act))
        return act

    def get_env(self, i):
        d = self.d
        hidden = np.empty(d * d, np.int8)
        syn = np.empty((d + 1) * (d + 1), np.int8)
        life = C.c_int32(); done = C.c_int32(); att = C.c_uint32()
        self.L.dqo_get_env(self.h, i, _p(hidden), _p(syn), C.byref(life), C.byref(done), C.byref(att))
        return dict(hidden=hidden.reshape(d, d), true_syndrome=syn.reshape(d + 1, d + 1),
                    lifetime=life.value, done=bool(done.value), attempts=att.value)

    def set_hidden(self, i, hidden):
        hidden = np.ascontiguousarray(hidden, dtype=np.int8).reshape(-1)
        self.L.dqo_set_hidden(self.h, i, _p(hidden))

    def syndrome_of(self, hidden):
        d = self.d
        hidden = np.ascontiguousarray(hidden, dtype=np.int8).reshape(-1)
        syn = np.empty((d + 1) * (d + 1), np.int8)
        label = C.c_int()
        self.L.dqo_syndrome_of(self.h, _p(hidden), _p(syn), C.byref(label))
        return syn.reshape(d + 1, d + 1), label.value

    def stab_order(self):
        a = np.empty(self.ns, np.int32); b = np.empty(self.ns, np.int32); t = np.empty(self.ns, np.int32)
        self.L.dqo_stab_order(self.h, _p(a), _p(b), _p(t))
        return a, b, t


def philox(c0, c1, c2, c3, k0, k1):
    out = np.empty(4, np.uint32)
    lib().dqo_philox(c0, c1, c2, c3, k0, k1, _p(out))
    return out
