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
