"""GPU parity of the packed host-buffer calls (dq_env_reset_host_packed / dq_env_step_host_packed) against the CPU oracle.
The same comparison runs on the CPU against the emulated kernel source in tests/test_env_emulated.py."""
import numpy as np
import pytest

from test_env_gpu import make_pair

pytestmark = pytest.mark.gpu


def test_packed_host_buffer_entry_points():
    """reset_host / step_host with packed=True: bit-packed observation rows over PCIe, expanded on the host."""
    from deepq_decoding_b200.envs import unpack_observations
    n = 100
    env, o = make_pair(5, "DP", False, 5, 0.02, n, seed=3)
    C = env.volume_depth + env.n_action_layers
    packed, legal = env.reset_host(packed=True); oobs, olegal = o.reset()
    assert np.array_equal(unpack_observations(packed, n, 5, C), oobs) and np.array_equal(legal, olegal)
    for t in range(40):
        acts = o.random_legal_actions(olegal, t)
        packed, rew, done, info = env.step_host(acts, packed=True)
        oobs, orew, odone, olife, olegal = o.step(acts, auto_reset=True)
        assert np.array_equal(unpack_observations(packed, n, 5, C), oobs), "packed obs t=%d" % t
        assert np.array_equal(rew, orew) and np.array_equal(done, odone.astype(bool))
        assert np.array_equal(info["lifetime"], olife) and np.array_equal(info["legal_mask"], olegal)
    env.close()
