"""Drives the UNMODIFIED reference environment (TEST INFRASTRUCTURE, build container only).

Imports /root/reference/cluster_scripts/d5_dp/{Environments,Function_Library}.py as
they lie (a 15-line `gym.spaces` stub stands in for the absent gym package) and feeds
them the shared counter-based random stream of DESIGN.md section 3 by rebinding the two
module globals through which the reference draws its noise inside reset()/step():

    Environments.generate_error           (called at Environments.py:162,221)
    Environments.generate_faulty_syndrome (called at Environments.py:166,225)

The reference's reset/step/legal-move/padding logic itself is untouched.  The referee
is any object with the duck-typed `.predict(x[1,(d+1)^2], batch_size=1, verbose=0)`
(Environments.py:144); `LutReferee` answers from the same table the CUDA path uses.

This module cannot travel to the GPU box (no /root/reference there): it is used here
to validate oracle/dq_oracle.c and to generate tests/golden/ fixtures.
"""
import os
import sys
import types

import numpy as np

REF_DIR = "/root/reference/cluster_scripts/d5_dp"


def available():
    return os.path.exists(os.path.join(REF_DIR, "Environments.py"))


def _install_gym_stub():
    if "gym" in sys.modules:
        return
    gym = types.ModuleType("gym")
    spaces = types.ModuleType("gym.spaces")

    class Box:
        def __init__(self, low, high, shape, dtype):
            self.low, self.high, self.shape, self.dtype = low, high, shape, dtype

    class Discrete:
        def __init__(self, n):
            self.n = n

    spaces.Box, spaces.Discrete = Box, Discrete
    gym.spaces = spaces
    sys.modules["gym"] = gym
    sys.modules["gym.spaces"] = spaces


_MODS = None


def reference_modules():
    """(Environments, Function_Library) of the reference, imported unmodified."""
    global _MODS
    if _MODS is None:
        if not available():
            raise RuntimeError("reference not present at " + REF_DIR)
        _install_gym_stub()
        sys.path.insert(0, REF_DIR)
        try:
            import Environments  # noqa
            import Function_Library  # noqa
        finally:
            sys.path.remove(REF_DIR)
        _MODS = (Environments, Function_Library)
    return _MODS


# ---- Philox4x32-10 in plain Python ints (independent of the C and CUDA copies) ----
_M0, _M1, _W0, _W1, _MASK = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85, 0xFFFFFFFF


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & _MASK, p1 & _MASK, ((p0 >> 32) ^ c3 ^ k1) & _MASK, p0 & _MASK
        k0 = (k0 + _W0) & _MASK
        k1 = (k1 + _W1) & _MASK
    return c0, c1, c2, c3


def threshold_u32(p):
    if not p > 0.0:
        return 0
    t = int(np.floor(p * 4294967296.0))
    return min(t, 0xFFFFFFFF)


class PhiloxNoise:
    """The noise stream of one environment (domain 0 of the contract)."""

    def __init__(self, d, volume_depth, error_model, seed, env_id):
        self.d, self.vd, self.model = d, volume_depth, error_model
        self.nq, self.ns = d * d, d * d - 1
        items = self.vd * (self.nq + self.ns)
        self.B = 32 * ((items + 127) // 128)
        self.k0, self.k1 = seed & _MASK, (seed >> 32) & _MASK
        self.env_id = env_id
        self.calls = 0          # completed (error, faulty) pairs; attempt = calls // vd
        self._cache = {}

    def word(self, attempt, i):
        b, w = i % self.B, i // self.B
        key = (attempt, b)
        if key not in self._cache:
            if len(self._cache) > 4096:
                self._cache.clear()
            self._cache[key] = philox4x32_10(self.env_id, attempt, b, 0, self.k0, self.k1)
        return self._cache[key][w]

    # drop-in for Function_Library.generate_error(d, p_phys, error_model)
    def generate_error(self, d, p_phys, error_model):
        attempt, j = divmod(self.calls, self.vd)
        T = threshold_u32(p_phys)
        T1, T2 = T // 3, (2 * T) // 3
        error = np.zeros((d, d), int)
        for q in range(d * d):
            u = self.word(attempt, j * self.nq + q)
            if error_model == "X":
                e = 1 if u < T else 0
            else:
                e = 1 if u < T1 else 2 if u < T2 else 3 if u < T else 0
            error[q // d, q % d] = e
        return error

    # drop-in for Function_Library.generate_faulty_syndrome(true_syndrome, p_meas)
    def generate_faulty_syndrome(self, true_syndrome, p_meas):
        attempt, j = divmod(self.calls, self.vd)
        Tm = threshold_u32(p_meas)
        g = true_syndrome.shape[0]
        faulty = np.zeros(np.shape(true_syndrome), int)
        nb = g // 2 - 1
        order = [(r, c) for r in range(1, g - 1) for c in range(1, g - 1)]
        order += [(0, 2 * x + 1) for x in range(nb)] + [(g - 1, 2 * x + 2) for x in range(nb)]
        order += [(2 * x + 2, 0) for x in range(nb)] + [(2 * x + 1, g - 1) for x in range(nb)]
        for k, (r, c) in enumerate(order):
            u = self.word(attempt, self.vd * self.nq + j * self.ns + k)
            faulty[r, c] = (1 - true_syndrome[r, c]) if u < Tm else true_syndrome[r, c]
        self.calls += 1
        return faulty


def stab_order(d):
    g = d + 1
    nb = g // 2 - 1
    order = [(r, c) for r in range(1, g - 1) for c in range(1, g - 1)]
    order += [(0, 2 * x + 1) for x in range(nb)] + [(g - 1, 2 * x + 2) for x in range(nb)]
    order += [(2 * x + 2, 0) for x in range(nb)] + [(2 * x + 1, g - 1) for x in range(nb)]
    return order


def joint_order(d):
    present = lambda a, b: not ((a == 0 and b % 2 == 0) or (a == d and b % 2 == 1) or
                                (b == 0 and a % 2 == 1) or (b == d and a % 2 == 0))
    order = [(a, b) for a in range(1, d) for b in range(d + 1) if present(a, b)]
    return order + [((0 if b % 2 else d), b) for b in range(1, d)]


class LutReferee:
    """`.predict` answering from the packed 2-bit referee tables (same bytes as the CUDA path)."""

    def __init__(self, d, error_model, mode, lut_a, lut_b=None):
        self.d, self.model, self.mode = d, error_model, mode
        self.lut_a = np.asarray(lut_a, np.uint8)
        self.lut_b = None if lut_b is None else np.asarray(lut_b, np.uint8)
        self.order = stab_order(d)
        self.jorder = joint_order(d)
        self.n_classes = 2 if error_model == "X" else 4

    @staticmethod
    def _get(lut, idx):
        return (int(lut[idx >> 2]) >> ((idx & 3) * 2)) & 3

    def classify(self, vec):
        g = self.d + 1
        allb = i3 = i1 = 0
        c3 = c1 = 0
        for k, (a, b) in enumerate(self.jorder):
            allb |= int(vec[a * g + b]) << k
        for k, (a, b) in enumerate(self.order):
            bit = int(vec[a * g + b])
            if (a + b) % 2 == 1:
                i3 |= bit << c3
                c3 += 1
            else:
                i1 |= bit << c1
                c1 += 1
        if self.mode == 0:
            return self._get(self.lut_a, allb)
        c = self._get(self.lut_a, i3) & 1
        if self.model == "DP" and self.lut_b is not None:
            c |= (self._get(self.lut_b, i1) & 1) << 1
        return c

    def predict(self, x, batch_size=1, verbose=0):
        out = np.zeros((len(x), self.n_classes), np.float32)
        for r, vec in enumerate(x):
            out[r, self.classify(vec)] = 1.0
        return out


class ReferenceEnv:
    """One unmodified reference env wired to the shared Philox stream.

    The two noise functions are module globals of `Environments`; since several
    envs with different streams may coexist, they are re-pointed at this env's
    stream around every reset()/step() call.
    """

    def __init__(self, d, error_model, use_Y, volume_depth, p_phys, p_meas, seed, env_id, referee):
        self.E, self.FL = reference_modules()
        self.noise = PhiloxNoise(d, volume_depth, error_model, seed, env_id)
        self.env = self.E.Surface_Code_Environment_Multi_Decoding_Cycles(
            d=d, p_phys=p_phys, p_meas=p_meas, error_model=error_model, use_Y=use_Y,
            volume_depth=volume_depth, static_decoder=referee)
        self.k0, self.k1, self.env_id = seed & _MASK, (seed >> 32) & _MASK, env_id

    def _bind(self):
        self.E.generate_error = self.noise.generate_error
        self.E.generate_faulty_syndrome = self.noise.generate_faulty_syndrome

    def _unbind(self):
        self.E.generate_error = self.FL.generate_error
        self.E.generate_faulty_syndrome = self.FL.generate_faulty_syndrome

    def reset(self):
        self._bind()
        try:
            return self.env.reset()
        finally:
            self._unbind()

    def step(self, action):
        self._bind()
        try:
            return self.env.step(int(action))
        finally:
            self._unbind()

    def random_legal_action(self, step):
        """Uniform over sorted(legal_actions), draw = domain-1 word of the contract."""
        u = philox4x32_10(self.env_id, step, 0, 1, self.k0, self.k1)[0]
        legal = sorted(self.env.legal_actions)
        return legal[(u * len(legal)) >> 32]

    def legal_mask(self):
        W = (self.env.num_actions + 63) // 64
        m = [0] * W
        for a in self.env.legal_actions:
            m[a >> 6] |= 1 << (a & 63)
        return np.array(m, dtype=np.uint64)
