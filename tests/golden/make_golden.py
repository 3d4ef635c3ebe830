"""Generates tests/golden/env_*.npz from the UNMODIFIED reference environment.

Run in the build container (needs /root/reference):   python tests/golden/make_golden.py

Each fixture is a trajectory of a few reference `Surface_Code_Environment_Multi_Decoding_Cycles`
instances (imported as shipped, see oracle/ref_harness.py) driven by the shared Philox noise stream
and the package's shipped referee table for that (d, error_model), with a recorded action sequence
(random-legal picks mixed with 25 % arbitrary, possibly illegal or repeated, actions).  A finished
episode is followed by `reset()` in the same step, which is the auto-reset contract of
`dq_env_step`: reward/done/lifetime describe the finished step, obs/legal the new episode.

The fixtures travel to the GPU box, where neither the reference nor this script can run:
tests/test_golden.py replays them through the CPU oracle (not gpu) and through the CUDA path (gpu).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_harness as RH  # noqa
from deepq_decoding_b200 import referee as REF  # noqa

#          name        d  model use_Y vd  p      steps envs
CONFIGS = [("d3_x",    3, "X",  False, 3, 0.05,  400, (0, 1, 2, 40)),     # BASELINE config 1
           ("d5_x",    5, "X",  False, 5, 0.03,  250, (0, 7, 33)),
           ("d5_dp",   5, "DP", False, 5, 0.03,  250, (0, 7, 33)),        # shipped-config shape, shipped referee
           ("d5_dpy",  5, "DP", True,  3, 0.04,  200, (3, 64)),           # the reference ctor's defaults
           ("d7_dp",   7, "DP", False, 7, 0.02,  120, (0, 31)),
           ("d3_dpy",  3, "DP", True,  2, 0.08,  300, (5, 6))]
SEED = 20260117


def make(name, d, model, use_Y, vd, p, steps, env_ids):
    referee = REF.shipped(d, model)
    rng = np.random.default_rng(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name)))
    envs = [RH.ReferenceEnv(d, model, use_Y, vd, p, p, SEED, i, referee) for i in env_ids]
    n = len(envs)
    A = envs[0].env.num_actions
    W = (A + 63) // 64
    obs0 = np.stack([e.reset().astype(np.uint8) for e in envs])
    legal0 = np.stack([e.legal_mask() for e in envs])
    life0 = np.array([e.env.lifetime for e in envs], np.int32)
    actions = np.zeros((steps, n), np.int32)
    obs = np.zeros((steps,) + obs0.shape, np.uint8)
    legal = np.zeros((steps, n, W), np.uint64)
    reward = np.zeros((steps, n), np.float32)
    done = np.zeros((steps, n), np.uint8)
    life = np.zeros((steps, n), np.int32)
    hidden = np.zeros((steps, n, d, d), np.int8)
    for t in range(steps):
        for k, e in enumerate(envs):
            a = e.random_legal_action(t) if rng.random() > 0.25 else int(rng.integers(0, A))
            actions[t, k] = a
            b, r, dn, _ = e.step(a)
            reward[t, k], done[t, k], life[t, k] = r, dn, e.env.lifetime
            if dn:
                b = e.reset()
            obs[t, k] = b.astype(np.uint8)
            legal[t, k] = e.legal_mask()
            hidden[t, k] = e.env.hidden_state.astype(np.int8)
    out = os.path.join(HERE, "env_%s.npz" % name)
    np.savez_compressed(out, d=d, model=model, use_Y=use_Y, vd=vd, p=p, seed=SEED, env_ids=np.array(env_ids),
                        obs0=np.packbits(obs0, axis=None), obs_shape=np.array(obs0.shape), legal0=legal0, life0=life0,
                        actions=actions, obs=np.packbits(obs, axis=None), legal=legal, reward=reward, done=done,
                        lifetime=life, hidden=hidden)
    print("%-8s steps=%d envs=%d done=%d heavy=%d  %d bytes" % (
        name, steps, n, int(done.sum()), int((np.diff(np.concatenate([life0[None], life]), axis=0) != 0).sum()),
        os.path.getsize(out)))


if __name__ == "__main__":
    for cfg in CONFIGS:
        make(*cfg)
