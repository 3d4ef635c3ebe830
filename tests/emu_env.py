"""TEST INFRASTRUCTURE: the environment's C ABI served by the CUDA kernel source executed on the CPU.

`tests/host/cuda_emu.h` + `tests/host/emu_build.py` compile `deepq_decoding_b200/csrc/dq_env.cu` itself (same templates, same
shared-memory struct, same warp primitives and barriers; only the `<<<...>>>` launch syntax and the dynamic shared-memory
declaration are rewritten, in a generated copy) into `tests/host/libdq_env_emu.so`; "device pointers" are numpy
buffers.  `EmuVecEnv` is the thin numpy caller the tests use to compare that build with the oracle when no
GPU is present, so a kernel change can be checked bit for bit before it ever reaches a B200.  Nothing under
`deepq_decoding_b200/` imports this, and nothing here is a fallback: the product raises without a GPU.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(ROOT, "deepq_decoding_b200", "csrc", "dq_env.cu")
SHIM = os.path.join(HERE, "host", "cuda_emu.h")
OUT = os.path.join(HERE, "host", "libdq_env_emu.so")
MODEL = {"X": 0, "DP": 1}
ROW_XB, ROW_ZB, ROW_META, ROW_ACT, ROW_SUM, ROW_BM = 0, 1, 2, 3, 6, 7


def build(extra_flags=(), out=OUT):
    """DQ_EMU_FLAGS="-DDQ_DEFER=1 ..." in the environment points the DEFAULT emulated library (every test that does not name a
    variant) at a tuning build, so the whole suite can be run against it before it is ever timed on a GPU."""
    import sys
    env_flags = os.environ.get("DQ_EMU_FLAGS", "").split()
    if env_flags and not extra_flags and out == OUT:
        import hashlib
        extra_flags = tuple(env_flags)
        out = OUT[:-3] + "_" + hashlib.md5(" ".join(env_flags).encode()).hexdigest()[:8] + ".so"
    sys.path.insert(0, os.path.join(HERE, "host"))
    import emu_build
    return emu_build.build(out, [SRC], extra_flags=tuple(extra_flags))


_LIB = {}


def lib(path=None):
    path = path or build()
    if path in _LIB:
        return _LIB[path]
    L = C.CDLL(path)
    vp, i32, i64, u64, f64 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_double
    L.dq_last_error.restype = C.c_char_p
    L.dq_env_create.argtypes = [C.POINTER(vp), i32, i32, i32, i32, f64, f64, i64, u64, i64, i32]
    L.dq_env_destroy.argtypes = [vp]
    L.dq_env_info.argtypes = [vp, i32, C.POINTER(i64)]
    L.dq_env_set_noise.argtypes = [vp, f64, f64]
    L.dq_env_set_max_attempts.argtypes = [vp, i32]
    L.dq_env_set_referee_lut.argtypes = [vp, i32, vp, i64, vp, i64]
    L.dq_env_reset.argtypes = [vp, vp, vp, vp]
    L.dq_env_step.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, vp]
    L.dq_env_step_random.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, vp]
    L.dq_env_rollout_random.argtypes = [vp, i32, vp, i32, i32, vp, vp, vp, vp, vp, i32, vp]
    L.dq_env_get_state.argtypes = [vp, vp, vp]
    L.dq_env_set_state.argtypes = [vp, vp, vp]
    L.dq_policy_random_legal.argtypes = [vp, vp, C.c_uint32, vp, vp]
    L.dq_policy_seek.argtypes = [vp, C.c_uint32, vp]
    L.dq_env_step_host.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32]
    L.dq_env_reset_host.argtypes = [vp, vp, vp]
    L.dq_env_step_host_begin.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32]
    L.dq_env_step_host_end.argtypes = [vp]
    L.dq_policy_random_legal_host.argtypes = [vp, vp, C.c_uint32, vp]
    L.dq_env_reset_host_packed.argtypes = [vp, vp, vp]
    L.dq_env_step_host_packed.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32]
    _LIB[path] = L
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class EmuVecEnv:
    def __init__(self, d, error_model, use_Y, volume_depth, p_phys, p_meas, n_envs, seed, env_id_base=0, lib_path=None):
        self.L = lib(lib_path)
        h = C.c_void_p()
        self._check(self.L.dq_env_create(C.byref(h), d, MODEL[error_model], int(use_Y), volume_depth, float(p_phys),
                                         float(p_meas), n_envs, seed, env_id_base, 0))
        self.h, self.n, self.d, self.vd = h, n_envs, d, volume_depth
        q = lambda w: self._info(w)
        self.A, self.Cn, self.H, self.W = q(0), q(1), q(2), q(3)
        self.state_rows, self.stride = q(4), q(5)
        self._luts = None

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("emulated ABI call failed (%d): %s" % (rc, self.L.dq_last_error().decode()))

    def _info(self, what):
        v = C.c_int64()
        self._check(self.L.dq_env_info(self.h, what, C.byref(v)))
        return int(v.value)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.dq_env_destroy(self.h)
            self.h = None

    def set_noise(self, p_phys, p_meas):
        self._check(self.L.dq_env_set_noise(self.h, float(p_phys), float(p_meas)))

    def set_max_attempts(self, n):
        self._check(self.L.dq_env_set_max_attempts(self.h, int(n)))

    def set_referee(self, mode, lut_a, lut_b=None):
        lut_a = np.ascontiguousarray(lut_a, dtype=np.uint8)
        lut_b = None if lut_b is None else np.ascontiguousarray(lut_b, dtype=np.uint8)
        self._luts = (lut_a, lut_b)
        self._check(self.L.dq_env_set_referee_lut(self.h, mode, _p(lut_a), lut_a.size, _p(lut_b), 0 if lut_b is None else lut_b.size))

    def _outs(self, rows=None):
        shp = (self.n,) if rows is None else (rows, self.n)
        return (np.zeros(shp, np.float32), np.zeros(shp, np.uint8), np.zeros(shp, np.int32),
                np.zeros(shp + (self.W,), np.uint64), np.zeros(shp, np.int32))

    def reset(self):
        obs = np.zeros((self.n, self.Cn, self.H, self.H), np.uint8)
        legal = np.zeros((self.n, self.W), np.uint64)
        self._check(self.L.dq_env_reset(self.h, _p(obs), _p(legal), None))
        return obs, legal

    def step(self, actions, auto_reset=True, want_obs=True):
        actions = np.ascontiguousarray(actions, dtype=np.int32)
        obs = np.zeros((self.n, self.Cn, self.H, self.H), np.uint8) if want_obs else None
        reward, done, life, legal, _ = self._outs()
        self._check(self.L.dq_env_step(self.h, _p(actions), _p(obs), _p(reward), _p(done), _p(life), _p(legal), int(auto_reset), None))
        return obs, reward, done, life, legal

    def step_random(self, auto_reset=True):
        obs = np.zeros((self.n, self.Cn, self.H, self.H), np.uint8)
        reward, done, life, legal, acts = self._outs()
        self._check(self.L.dq_env_step_random(self.h, _p(obs), _p(reward), _p(done), _p(life), _p(legal), _p(acts), int(auto_reset), None))
        return obs, reward, done, life, legal, acts

    def rollout_random(self, n_steps, slots, first_slot=0, auto_reset=True):
        ring = np.zeros((slots, self.n, self.Cn, self.H, self.H), np.uint8)
        reward, done, life, legal, acts = self._outs(n_steps)
        self._check(self.L.dq_env_rollout_random(self.h, n_steps, _p(ring), slots, first_slot, _p(reward), _p(done), _p(life),
                                                 _p(legal), _p(acts), int(auto_reset), None))
        return ring, reward, done, life, legal, acts

    def random_legal_actions(self, legal, step):
        legal = np.ascontiguousarray(legal, dtype=np.uint64)
        acts = np.zeros(self.n, np.int32)
        self._check(self.L.dq_policy_random_legal(self.h, _p(legal), step, _p(acts), None))
        return acts

    def policy_seek(self, step):
        self._check(self.L.dq_policy_seek(self.h, step, None))

    def state_words(self):
        w = np.zeros((self.state_rows, self.stride), np.uint64)
        self._check(self.L.dq_env_get_state(self.h, _p(w), None))
        return w

    def step_host(self, actions, auto_reset=True):
        actions = np.ascontiguousarray(actions, dtype=np.int32)
        obs = np.zeros((self.n, self.Cn, self.H, self.H), np.uint8)
        reward, done, life, legal, _ = self._outs()
        self._check(self.L.dq_env_step_host(self.h, _p(actions), _p(obs), _p(reward), _p(done), _p(life), _p(legal), int(auto_reset)))
        return obs, reward, done, life, legal

    def step_host_begin(self, actions, auto_reset=True):
        actions = np.ascontiguousarray(actions, dtype=np.int32)
        obs = np.zeros((self.n, self.Cn, self.H, self.H), np.uint8)
        reward, done, life, legal, _ = self._outs()
        self._check(self.L.dq_env_step_host_begin(self.h, _p(actions), _p(obs), _p(reward), _p(done), _p(life), _p(legal), int(auto_reset)))
        self._pending_actions, self._pending = actions, (obs, reward, done, life, legal)      # kept alive until step_host_end (a refused call keeps the earlier ones)

    def step_host_end(self):
        self._check(self.L.dq_env_step_host_end(self.h))
        return self._pending

    def random_legal_actions_host(self, legal, step):
        legal = np.ascontiguousarray(legal, dtype=np.uint64)
        acts = np.zeros(self.n, np.int32)
        self._check(self.L.dq_policy_random_legal_host(self.h, _p(legal), step, _p(acts)))
        return acts

    def step_host_packed(self, actions, auto_reset=True):
        actions = np.ascontiguousarray(actions, dtype=np.int32)
        packed = np.zeros((self.state_rows - ROW_BM, self.stride), np.uint64)
        reward, done, life, legal, _ = self._outs()
        self._check(self.L.dq_env_step_host_packed(self.h, _p(actions), _p(packed), _p(reward), _p(done), _p(life), _p(legal), int(auto_reset)))
        return packed, reward, done, life, legal

    def reset_host(self):
        obs = np.zeros((self.n, self.Cn, self.H, self.H), np.uint8)
        legal = np.zeros((self.n, self.W), np.uint64)
        self._check(self.L.dq_env_reset_host(self.h, _p(obs), _p(legal)))
        return obs, legal

    def reset_host_packed(self):
        packed = np.zeros((self.state_rows - ROW_BM, self.stride), np.uint64)
        legal = np.zeros((self.n, self.W), np.uint64)
        self._check(self.L.dq_env_reset_host_packed(self.h, _p(packed), _p(legal)))
        return packed, legal
