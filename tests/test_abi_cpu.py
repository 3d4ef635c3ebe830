"""CPU-only checks of the drop-in boundary: the shared library loads, exports every symbol that
include/dq_decoding.h declares, and fails loudly (no CPU fallback) when asked to compute without a GPU."""
import ctypes as C
import os
import re

import pytest

from deepq_decoding_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "dq_decoding.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dq_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == _lib.exported_symbols()


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    for name in declared_symbols():
        assert hasattr(L, name), name
    assert L.dq_version() >= 100


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = _lib.lib()
    h = C.c_void_p()
    rc = L.dq_env_create(C.byref(h), 5, 1, 0, 5, 0.007, 0.007, 64, 0, 0, 0)
    assert rc == _lib.ECUDA and not h.value
    assert L.dq_last_error()
    from deepq_decoding_b200.envs import VecSurfaceCodeEnv
    with pytest.raises(Exception):
        VecSurfaceCodeEnv(5, 0.007, 0.007, "DP", False, 5, None, n_envs=4)


def test_argument_validation_precedes_device_use():
    L = _lib.lib()
    h = C.c_void_p()
    assert L.dq_env_create(C.byref(h), 4, 1, 0, 5, 0.01, 0.01, 8, 0, 0, 0) == _lib.EINVAL
    assert b"d must be" in L.dq_last_error()
    assert L.dq_env_create(C.byref(h), 5, 7, 0, 5, 0.01, 0.01, 8, 0, 0, 0) == _lib.EINVAL
    assert L.dq_env_create(C.byref(h), 5, 1, 0, 0, 0.01, 0.01, 8, 0, 0, 0) == _lib.EINVAL
    assert L.dq_env_create(C.byref(h), 5, 1, 0, 5, 0.01, 0.01, 0, 0, 0, 0) == _lib.EINVAL
    assert L.dq_env_step(None, None, None, None, None, None, None, 1, None) == _lib.EINVAL


_UNPACK_CHILD = r"""
import ctypes as C, sys
import numpy as np
from deepq_decoding_b200 import _lib
from deepq_decoding_b200.envs import unpack_observations
L = _lib.lib()
rng = np.random.default_rng(5)
vp = lambda a: C.c_void_p(a.ctypes.data)
for d, ch, n, stride in ((3, 4, 1, 32), (5, 7, 1000, 1024), (5, 6, 4097, 4128), (7, 9, 333, 352)):
    side = 2 * d + 1
    P, PW = side * side, (side * side + 63) // 64
    rows = rng.integers(0, 2 ** 63, size=(ch * PW, stride), dtype=np.int64).view(np.uint64)
    rows[PW - 1::PW] &= np.uint64((1 << (P - 64 * (PW - 1))) - 1)            # the kernel keeps bits >= P of a bitmap zero
    obs = np.full((n + 1, ch, side, side), 7, dtype=np.uint8)                  # one lattice of slack: nothing may be written past n
    assert L.dq_unpack_observations_host(vp(rows), stride, n, d, ch, vp(obs)) == 0
    assert np.array_equal(obs[:n], unpack_observations(rows, n, d, ch)), (d, ch, n)
    assert (obs[n] == 7).all(), "wrote past the last lattice"
assert L.dq_unpack_observations_host(None, 1, 1, 5, 7, None) == _lib.EINVAL
assert L.dq_unpack_observations_host(vp(rows), 4, 8, 5, 7, vp(obs)) == _lib.EINVAL      # stride < n
print("ok")
"""


@pytest.mark.parametrize("isa", ["default", "no_avx512", "portable"])
def test_host_unpack_matches_the_numpy_helper(isa):
    """dq_unpack_observations_host (the expansion inside dq_env_step_host, on its own) against envs.unpack_observations, for each
    instruction-set path of the library (chosen once per process, hence the child) and with several host threads."""
    import subprocess
    import sys
    env = dict(os.environ, PYTHONPATH=ROOT, DQ_HOST_THREADS="4")
    if isa == "no_avx512":
        env["DQ_HOST_NO_AVX512"] = "1"
    if isa == "portable":
        env["DQ_HOST_NO_AVX2"] = "1"
    r = subprocess.run([sys.executable, "-c", _UNPACK_CHILD], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]
