"""Replays the reference-generated golden trajectories (tests/golden/make_golden.py).

CPU:  the oracle must reproduce every recorded output bit for bit (this is the oracle's pin that
      travels with the repo).
GPU:  the CUDA path, called through the C ABI, must do the same.
"""
import glob
import os

import numpy as np
import pytest

from oracle import oracle as O
from deepq_decoding_b200 import referee as REF

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden", "env_*.npz")))


def load(path):
    z = np.load(path)
    g = {k: z[k] for k in z.files}
    shape = tuple(g["obs_shape"])
    T = g["actions"].shape[0]
    g["obs0"] = np.unpackbits(g["obs0"])[:int(np.prod(shape))].reshape(shape)
    g["obs"] = np.unpackbits(g["obs"])[:T * int(np.prod(shape))].reshape((T,) + shape)
    for k in ("d", "vd"):
        g[k] = int(g[k])
    g["model"], g["use_Y"], g["p"], g["seed"] = str(g["model"]), bool(g["use_Y"]), float(g["p"]), int(g["seed"])
    return g


def test_fixtures_present():
    assert len(FILES) >= 6


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_oracle_reproduces_reference(path):
    g = load(path)
    ids = [int(i) for i in g["env_ids"]]
    n = max(ids) + 1
    o = O.OracleVecEnv(g["d"], g["model"], g["use_Y"], g["vd"], g["p"], g["p"], n, g["seed"])
    ref = REF.shipped(g["d"], g["model"])
    o.set_referee(ref.mode, ref.lut_a, ref.lut_b)
    obs, legal = o.reset()
    assert np.array_equal(obs[ids], g["obs0"])
    assert np.array_equal(legal[ids], g["legal0"])
    acts = np.full(n, o.A - 1, np.int32)
    for t in range(g["actions"].shape[0]):
        acts[:] = o.A - 1
        acts[ids] = g["actions"][t]
        obs, rew, done, life, legal = o.step(acts, auto_reset=True)
        assert np.array_equal(obs[ids], g["obs"][t]), t
        assert np.array_equal(rew[ids], g["reward"][t]), t
        assert np.array_equal(done[ids], g["done"][t]), t
        assert np.array_equal(life[ids], g["lifetime"][t]), t
        assert np.array_equal(legal[ids], g["legal"][t]), t
        for k, i in enumerate(ids):
            assert np.array_equal(o.get_env(i)["hidden"], g["hidden"][t, k])
    assert g["done"].sum() > 0


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_cuda_reproduces_reference(path):
    import torch
    from deepq_decoding_b200.envs import VecSurfaceCodeEnv
    g = load(path)
    ids = [int(i) for i in g["env_ids"]]
    n = max(ids) + 1
    env = VecSurfaceCodeEnv(g["d"], g["p"], g["p"], g["model"], g["use_Y"], g["vd"], None, n_envs=n, seed=g["seed"])
    obs = env.reset().cpu().numpy()
    assert np.array_equal(obs[ids], g["obs0"])
    assert np.array_equal(env.legal_mask.cpu().numpy().view(np.uint64)[ids], g["legal0"])
    acts = np.full(n, env.num_actions - 1, np.int32)
    for t in range(g["actions"].shape[0]):
        acts[:] = env.num_actions - 1
        acts[ids] = g["actions"][t]
        obs, rew, done, info = env.step(torch.from_numpy(acts).cuda())
        assert np.array_equal(obs.cpu().numpy()[ids], g["obs"][t]), t
        assert np.array_equal(rew.cpu().numpy()[ids], g["reward"][t]), t
        assert np.array_equal(done.cpu().numpy()[ids], g["done"][t].astype(bool)), t
        assert np.array_equal(info["lifetime"].cpu().numpy()[ids], g["lifetime"][t]), t
        assert np.array_equal(info["legal_mask"].cpu().numpy().view(np.uint64)[ids], g["legal"][t]), t
        if t % 25 == 0:
            st = env.decode_state()
            for k, i in enumerate(ids):
                assert np.array_equal(st[i]["hidden_state"], g["hidden"][t, k])
    env.close()
