"""Pins of the shipped referee lookup tables (deepq_decoding_b200/data/referee_d5_{X,DP}.lut, what the env kernel gathers from)
against the reference's Keras referee MLPs (example_notebooks/referee_decoders/nn_d5_*_p5; call site Environments.py:144-150).

tests/golden/referee_pins.npz (tests/golden/make_golden_pins.py) holds the MLP's argmax, evaluated in float64 numpy from the HDF5
weights, for ALL 2^12 syndromes bit-flip noise can produce and for 200 000 random depolarising syndromes: the tables must answer
the same class for every one of them.  With /root/reference present the MLPs are re-evaluated as well (fp64 and fp32).
"""
import os

import numpy as np
import pytest

from deepq_decoding_b200 import referee as R

HERE = os.path.dirname(os.path.abspath(__file__))
PINS = np.load(os.path.join(HERE, "golden", "referee_pins.npz"))
D, G = 5, 6


def table_classes(ref, order, index):
    """The class the packed table holds for syndromes given as bit patterns over `order` (vectorised RefereeLUT.classify)."""
    pos = {tuple(ab): k for k, ab in enumerate(order)}
    index = index.astype(np.int64)
    bit = lambda ab: ((index >> pos[ab]) & 1) if ab in pos else np.zeros_like(index)
    if ref.mode == R.JOINT:
        idx = np.zeros_like(index)
        for k, ab in enumerate(R.joint_order(D)):
            idx |= bit(ab) << k
        return (ref.lut_a[idx >> 2] >> ((idx & 3) * 2)) & 3
    out = np.zeros_like(index)
    for odd, lut, shift in ((1, ref.lut_a, 0), (0, ref.lut_b, 1)):
        if lut is None:
            continue
        idx = np.zeros_like(index)
        for k, ab in enumerate(R.type_order(D, odd)):
            idx |= bit(ab) << k
        out |= ((lut[idx >> 2] >> ((idx & 3) * 2)) & 1) << shift
    return out


@pytest.mark.parametrize("model", ["X", "DP"])
def test_shipped_table_answers_like_the_keras_referee(model):
    ref = R.shipped(D, model)
    order = [tuple(int(v) for v in ab) for ab in PINS[model + "/order"]]
    index, want = PINS[model + "/index"], PINS[model + "/class"]
    if model == "X":
        assert len(index) == 1 << 12 and np.array_equal(index, np.arange(1 << 12)), "exhaustive over the reachable syndromes"
    else:
        assert len(index) >= 100000
    got = table_classes(ref, order, index)
    assert np.array_equal(got, want), "%d of %d syndromes classified differently" % (int((got != want).sum()), len(want))
    # spot-check the scalar path the unmodified reference env calls (.predict on the flattened (d+1)^2 syndrome)
    for i in range(0, len(index), max(1, len(index) // 50)):
        vec = np.zeros(G * G, int)
        for k, (a, b) in enumerate(order):
            vec[a * G + b] = (int(index[i]) >> k) & 1
        assert int(np.argmax(ref.predict(vec[None])[0])) == int(want[i])


@pytest.mark.reference
@pytest.mark.parametrize("model,fn", [("X", "nn_d5_X_p5"), ("DP", "nn_d5_DP_p5")])
def test_fixture_and_table_against_the_hdf5_mlp(model, fn):
    """Re-evaluates the shipped HDF5 MLP here: float64 reproduces the fixture; float32 (what Keras predict computes in) agrees on
    every pinned syndrome too, so the canonical fp64 table is also what the reference's own arithmetic decides."""
    path = os.path.join("/root/reference/example_notebooks/referee_decoders", fn)
    layers = R.load_keras_mlp(path)
    assert [k.shape for k, _ in layers] == [(36, 1000), (1000, 500), (500, 250), (250, 50), (50, 2 if model == "X" else 4)]
    order = [tuple(int(v) for v in ab) for ab in PINS[model + "/order"]]
    index, want = PINS[model + "/index"][:40000], PINS[model + "/class"][:40000]
    x = np.zeros((len(index), G * G))
    for k, (a, b) in enumerate(order):
        x[:, a * G + b] = (index.astype(np.int64) >> k) & 1
    for dt in (np.float64, np.float32):
        h = x.astype(dt)
        for i, (W, b) in enumerate(layers):
            h = h @ W.astype(dt) + b.astype(dt)
            if i + 1 < len(layers):
                h = np.maximum(h, 0)
        assert np.array_equal(np.argmax(h, axis=1), want), dt
