"""Pins of host-visible conventions against the UNMODIFIED reference functions (build container only: /root/reference).

  * draw order and present-stabilizer set of generate_faulty_syndrome (Function_Library.py:186-233, CS copy :176-223): the
    RNG contract's measurement draws follow it; checked by letting only the k-th `np.random.rand()` fire;
  * class split of generate_DP_error (Function_Library.py:96-113): rand() < p, then randint(1,4) -- X, Y, Z with p/3 each; the
    contract replaces the two draws by one word against T/3, 2T/3, T;
  * indicate_identity / generate_identity_indicator / padding_syndrome / padding_actions (Environments.py:273-324, 374-385):
    the adapter's host helpers against the reference methods on random inputs.
"""
import types

import numpy as np
import pytest

from oracle import ref_harness as RH
from oracle import oracle as O
from deepq_decoding_b200 import referee as R

pytestmark = pytest.mark.reference


@pytest.mark.parametrize("d", [3, 5, 7])
def test_faulty_syndrome_draw_order_and_present_set(d):
    E, FL = RH.reference_modules()
    g = d + 1
    order_product = R.stabilizer_order(d)
    order_harness = RH.stab_order(d)
    o = O.OracleVecEnv(d, "DP", False, d, 0.01, 0.01, 1, 0)
    oa, ob, _ = o.stab_order()
    order_oracle = list(zip([int(x) for x in oa], [int(x) for x in ob]))
    rng = np.random.default_rng(d)
    true = rng.integers(0, 2, size=(g, g))                  # junk on the absent cells as well: they must come back 0
    real_rand = FL.np.random.rand
    try:
        calls = {"n": 0, "fire": -1}

        def fake_rand():
            k = calls["n"]
            calls["n"] += 1
            return 0.0 if k == calls["fire"] else 1.0

        FL.np.random.rand = fake_rand
        calls["n"], calls["fire"] = 0, -1
        plain = FL.generate_faulty_syndrome(true, 0.5)
        ndraws = calls["n"]
        assert ndraws == d * d - 1, "one draw per present stabilizer"
        present = {(a, b) for a in range(g) for b in range(g) if R.plaquette_present(d, a, b)}
        assert len(present) == d * d - 1
        for a in range(g):
            for b in range(g):
                assert plain[a, b] == (true[a, b] if (a, b) in present else 0)
        flipped = []
        for k in range(ndraws):
            calls["n"], calls["fire"] = 0, k
            out = FL.generate_faulty_syndrome(true, 0.5)
            diff = np.argwhere(out != plain)
            assert len(diff) == 1, "draw %d flips exactly one stabilizer" % k
            flipped.append((int(diff[0][0]), int(diff[0][1])))
    finally:
        FL.np.random.rand = real_rand
    assert flipped == order_product == order_harness == order_oracle
    assert set(flipped) == present


def test_depolarising_sampler_class_split():
    """Reference: P(X) = P(Y) = P(Z) = p/3.  Contract: thresholds T1 = T//3, T2 = 2T//3, T = floor(p 2^32) on one uniform word."""
    E, FL = RH.reference_modules()
    for p in (0.007, 0.011, 0.3):
        T = RH.threshold_u32(p)
        T1, T2 = T // 3, (2 * T) // 3
        probs = np.array([T1, T2 - T1, T - T2]) / 2.0 ** 32
        assert np.all(np.abs(probs - p / 3) < 2.0 ** -31), "contract thresholds split p into thirds up to 2^-32"
    p, d, calls = 0.3, 7, 3000
    state = np.random.get_state()
    np.random.seed(12345)
    try:
        counts = np.zeros(4, np.int64)
        for _ in range(calls):
            counts += np.bincount(FL.generate_DP_error(d, p).reshape(-1), minlength=4)
    finally:
        np.random.set_state(state)
    n = calls * d * d
    T = RH.threshold_u32(p)
    expect = np.array([2 ** 32 - T, T // 3, (2 * T) // 3 - T // 3, T - (2 * T) // 3]) / 2.0 ** 32 * n
    chi2 = float(((counts - expect) ** 2 / expect).sum())
    assert chi2 < 21.1, "reference sampler vs contract class probabilities: chi2 = %.1f (3 dof, p < 1e-4)" % chi2
    # and the contract's own sampler (the harness feeds it to the reference env) draws the same classes
    noise = RH.PhiloxNoise(d, d, "DP", seed=99, env_id=1)
    c2 = np.zeros(4, np.int64)
    for k in range(600):
        c2 += np.bincount(noise.generate_error(d, p, "DP").reshape(-1), minlength=4)
        noise.calls += 1
    e2 = expect / n * c2.sum()
    assert float(((c2 - e2) ** 2 / e2).sum()) < 21.1
    # bit-flip model: rand() < p -> X
    np.random.seed(7)
    cx = sum(int(FL.generate_X_error(d, p).sum()) for _ in range(1000))
    assert abs(cx - p * 1000 * d * d) < 5 * np.sqrt(p * (1 - p) * 1000 * d * d)
    np.random.set_state(state)


@pytest.mark.parametrize("d,model,use_Y,vd", [(3, "X", False, 3), (5, "DP", False, 5), (5, "DP", True, 3), (7, "DP", False, 7)])
def test_adapter_host_helpers_match_reference_methods(d, model, use_Y, vd):
    E, FL = RH.reference_modules()
    from deepq_decoding_b200.envs import Surface_Code_Environment_Multi_Decoding_Cycles as Adapter
    ref = E.Surface_Code_Environment_Multi_Decoding_Cycles(d=d, p_phys=0.01, p_meas=0.01, error_model=model, use_Y=use_Y,
                                                           volume_depth=vd, static_decoder=None)
    layers = ref.n_action_layers
    me = types.SimpleNamespace(d=d, volume_depth=vd, n_action_layers=layers)      # the helpers only read these attributes
    me.identity_indicator = Adapter.generate_identity_indicator(me, d)
    assert np.array_equal(me.identity_indicator, ref.generate_identity_indicator(d))
    assert np.array_equal(me.identity_indicator, ref.identity_indicator)
    rng = np.random.default_rng(d + vd)
    for _ in range(5):
        board = rng.integers(0, 2, size=(vd + layers, 2 * d + 1, 2 * d + 1))
        assert np.array_equal(Adapter.indicate_identity(me, board.copy()), ref.indicate_identity(board.copy()))
        syn = rng.integers(0, 2, size=(d + 1, d + 1))
        assert np.array_equal(Adapter.padding_syndrome(me, syn), ref.padding_syndrome(syn))
        acts = rng.integers(0, 2, size=d * d)
        assert np.array_equal(Adapter.padding_actions(me, acts), ref.padding_actions(acts))
