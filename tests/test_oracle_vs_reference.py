"""Pins the CPU oracle (oracle/dq_oracle.c) against the UNMODIFIED reference classes.

Both sides consume the same Philox stream (oracle/ref_harness.py rebinds the two
noise functions of the reference; its reset/step logic runs as shipped).  Compared
after every step: board_state (all cells), reward, done, lifetime, legal_actions,
hidden_state, current_true_syndrome.  30 % of the actions are arbitrary (possibly
illegal / repeated) to exercise the identity-by-repetition path.  Runs only where
/root/reference exists; tests/golden/ carries the same trajectories to the GPU box.
"""
import numpy as np
import pytest

from oracle import oracle as O
from oracle import ref_harness as R

pytestmark = pytest.mark.reference


def random_referee(rng, o, model, d):
    if model == "X" or d == 7:
        la = rng.integers(0, 256, size=(1 << o.n3) // 4 + 1, dtype=np.uint8) & 0x55
        lb = rng.integers(0, 256, size=(1 << o.n1) // 4 + 1, dtype=np.uint8) & 0x55
        return 1, la, lb
    return 0, rng.integers(0, 256, size=(1 << o.ns) // 4 + 1, dtype=np.uint8), None


def run(d, model, use_Y, vd, p, steps, seed=1234, env_ids=(0, 5), illegal_frac=0.3):
    rng = np.random.default_rng(d * 100 + vd)
    o = O.OracleVecEnv(d, model, use_Y, vd, p, p, max(env_ids) + 1, seed)
    mode, la, lb = random_referee(rng, o, model, d)
    o.set_referee(mode, la, lb)
    refs = {i: R.ReferenceEnv(d, model, use_Y, vd, p, p, seed, i, R.LutReferee(d, model, mode, la, lb))
            for i in env_ids}

    def check_reset():
        obs, legal = o.reset()
        for i, r in refs.items():
            b = r.reset()
            assert np.array_equal(b.astype(np.uint8), obs[i])
            assert np.array_equal(r.legal_mask(), legal[i])
            assert r.env.lifetime == o.get_env(i)["lifetime"]
        return legal

    legal = check_reset()
    ndone = 0
    for t in range(steps):
        acts = o.random_legal_actions(legal, t)
        arb = rng.random(o.n) < illegal_frac
        acts[arb] = rng.integers(0, o.A, size=arb.sum())
        for i, r in refs.items():
            if not arb[i]:
                assert r.random_legal_action(t) == acts[i]
        obs, rew, done, life, legal = o.step(acts, auto_reset=False)
        for i, r in refs.items():
            b, rr, dd, _ = r.step(acts[i])
            assert np.array_equal(b.astype(np.uint8), obs[i]), f"obs t={t}"
            assert rr == rew[i], f"reward t={t}"
            assert bool(dd) == bool(done[i]), f"done t={t}"
            assert r.env.lifetime == life[i]
            assert np.array_equal(r.legal_mask(), legal[i]), f"legal t={t}"
            st = o.get_env(i)
            assert np.array_equal(st["hidden"], r.env.hidden_state.astype(int))
            assert np.array_equal(st["true_syndrome"], r.env.current_true_syndrome)
            ndone += bool(dd)
        if done.any():
            legal = check_reset()
    return ndone


@pytest.mark.parametrize("cfg", [
    (3, "X", False, 3, 0.05, 1500),     # BASELINE config 1
    (5, "X", False, 5, 0.02, 600),
    (5, "DP", False, 5, 0.02, 600),     # shipped-config shape (use_Y False, vd 5)
    (5, "DP", True, 3, 0.03, 400),      # ctor defaults of the reference env
    (7, "DP", False, 7, 0.011, 200),
    (3, "DP", True, 2, 0.08, 500),
])
def test_trajectory_parity(cfg):
    d, model, use_Y, vd, p, steps = cfg
    ndone = run(d, model, use_Y, vd, p, steps)
    assert ndone > 0        # the terminal branch was exercised


def test_philox_matches_python():
    rng = np.random.default_rng(0)
    for _ in range(50):
        c = [int(x) for x in rng.integers(0, 2**32, size=6)]
        assert tuple(int(x) for x in O.philox(*c)) == R.philox4x32_10(*c)
    # Random123 known-answer vectors for philox4x32-10
    assert R.philox4x32_10(0, 0, 0, 0, 0, 0) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert R.philox4x32_10(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff) == \
        (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert R.philox4x32_10(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)
