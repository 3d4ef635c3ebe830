"""CPU checks of the device bit-board helpers (dq_lattice.cuh compiled for the host) against the oracle.

The CUDA kernel is built from these helpers; checking each one here means a GPU run only has to
prove the warp-level choreography, not the lattice arithmetic.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O
from deepq_decoding_b200 import referee as R

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def hb():
    so = os.path.join(HERE, "host", "libhost_bits.so")
    src = os.path.join(HERE, "host", "host_bits_check.cpp")
    hdr = os.path.join(ROOT, "deepq_decoding_b200", "csrc", "dq_lattice.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, src])
    L = C.CDLL(so)
    u64, u32, i = C.c_uint64, C.c_uint32, C.c_int
    for name, res, args in [("hb_true_syndrome", u64, [i, u64, u64]), ("hb_label", i, [i, u64, u64]),
                            ("hb_q_c2g", u64, [i, u64]), ("hb_q_g2c", u64, [i, u64]), ("hb_s_c2g", u64, [i, u64]),
                            ("hb_s_g2c", u64, [i, u64]), ("hb_type_index", u32, [i, i, u64]), ("hb_joint_index", u32, [i, u64]),
                            ("hb_adjacent", u64, [i, u64]), ("hb_neighbours", u64, [i, u64]),
                            ("hb_syn_layer", None, [i, u64, C.c_void_p]), ("hb_act_layer", None, [i, u64, C.c_void_p]),
                            ("hb_philox", None, [u32] * 6 + [C.c_void_p]), ("hb_extract_bits", u64, [C.c_void_p, i, i]),
                            ("hb_select64", i, [u64, i]), ("hb_masks", u64, [i, i])]:
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    return L


def grid_word(arr):
    """2-D 0/1 array -> uint64 with bit r*(d+1)+c, d+1 = stride of the plaquette grid."""
    rows, cols = arr.shape
    g = max(rows, cols) if rows == cols and False else None
    return arr, g


def to_word(arr, g):
    w = 0
    for r in range(arr.shape[0]):
        for c in range(arr.shape[1]):
            if arr[r, c]:
                w |= 1 << (r * g + c)
    return w


def from_word(w, g, rows, cols):
    return np.array([[(w >> (r * g + c)) & 1 for c in range(cols)] for r in range(rows)])


@pytest.mark.parametrize("d", [3, 5, 7])
def test_syndrome_label_and_orders(hb, d):
    rng = np.random.default_rng(d)
    o = O.OracleVecEnv(d, "DP", False, d, 0.01, 0.01, 1, 0)
    g = d + 1
    sa, sb, st = o.stab_order()
    assert [(int(a), int(b)) for a, b in zip(sa, sb)] == R.stabilizer_order(d)
    t1 = sum(1 << (int(a) * g + int(b)) for a, b, t in zip(sa, sb, st) if t == 1)
    t3 = sum(1 << (int(a) * g + int(b)) for a, b, t in zip(sa, sb, st) if t == 3)
    assert hb.hb_masks(d, 1) == t1 and hb.hb_masks(d, 2) == t3
    for _ in range(300):
        hidden = rng.integers(0, 4, size=(d, d)) * (rng.random((d, d)) < rng.choice([0.1, 0.5, 1.0]))
        xb = to_word((hidden == 1) | (hidden == 2), g)
        zb = to_word((hidden == 2) | (hidden == 3), g)
        syn, label = o.syndrome_of(hidden)
        s = hb.hb_true_syndrome(d, xb, zb)
        assert np.array_equal(from_word(s, g, g, g), syn)
        assert hb.hb_label(d, xb, zb) == label
        # draw-order compaction and its inverse; per-type indices
        c = hb.hb_s_g2c(d, s)
        assert c == sum(int(syn[a, b]) << k for k, (a, b) in enumerate(R.stabilizer_order(d)))
        assert hb.hb_s_c2g(d, c) == s
        if d <= 5:
            assert hb.hb_joint_index(d, s) == sum(int(syn[a, b]) << k for k, (a, b) in enumerate(R.joint_order(d)))
        for odd in (0, 1):
            assert hb.hb_type_index(d, odd, s) == sum(int(syn[a, b]) << k for k, (a, b) in enumerate(R.type_order(d, odd)))
        q = int(rng.integers(0, 1 << (d * d)))
        assert hb.hb_q_g2c(d, hb.hb_q_c2g(d, q)) == q
        assert hb.hb_q_c2g(d, q) == sum(((q >> (r * d + cc)) & 1) << (r * g + cc) for r in range(d) for cc in range(d))


@pytest.mark.parametrize("d", [3, 5, 7])
def test_legal_move_sets(hb, d):
    rng = np.random.default_rng(10 + d)
    g = d + 1
    for _ in range(200):
        summed = np.zeros((g, g), int)
        for a, b in R.stabilizer_order(d):
            summed[a, b] = rng.random() < 0.15
        adj = hb.hb_adjacent(d, to_word(summed, g))
        exp = np.zeros((d, d), int)
        for r in range(d):
            for c in range(d):
                exp[r, c] = summed[r, c] | summed[r, c + 1] | summed[r + 1, c] | summed[r + 1, c + 1]
        assert np.array_equal(from_word(adj, g, d, d), exp)
        acted = (rng.random((d, d)) < 0.15).astype(int)
        n8 = hb.hb_neighbours(d, to_word(acted, g))
        exp = np.zeros((d, d), int)
        for r in range(d):
            for c in range(d):
                if acted[r, c]:
                    for dr in (-1, 0, 1):
                        for dc in (-1, 0, 1):
                            if (dr or dc) and 0 <= r + dr < d and 0 <= c + dc < d:
                                exp[r + dr, c + dc] = 1
        assert np.array_equal(from_word(n8, g, d, d), exp)


@pytest.mark.parametrize("d", [3, 5, 7])
def test_observation_layers(hb, d):
    """Layer bitmaps == the oracle's padded board (Environments.py:273-314) cell for cell."""
    rng = np.random.default_rng(20 + d)
    g, H = d + 1, 2 * d + 1
    nw = (H * H + 63) // 64
    out = np.zeros(nw, np.uint64)
    for _ in range(100):
        f = np.zeros((g, g), int)
        for a, b in R.stabilizer_order(d):
            f[a, b] = rng.random() < 0.3
        hb.hb_syn_layer(d, to_word(f, g), out.ctypes.data_as(C.c_void_p))
        bits = np.array([(int(out[i >> 6]) >> (i & 63)) & 1 for i in range(H * H)]).reshape(H, H)
        exp = np.zeros((H, H), int)
        for x in range(H):
            for y in range(H):
                if (x in (0, 2 * d) and y % 2 == 1) or (y in (0, 2 * d) and x % 2 == 1):
                    exp[x, y] = 1
                if x % 2 == 0 and y % 2 == 0:
                    exp[x, y] = f[x // 2, y // 2]
                elif x % 2 == 1 and y % 2 == 1 and (x + y) % 4 == 0:
                    exp[x, y] = 1
        assert np.array_equal(bits, exp)
        assert all(int(out[i >> 6]) >> (i & 63) == 0 for i in [H * H] if (H * H) & 63)
        act = (rng.random((d, d)) < 0.3).astype(int)
        hb.hb_act_layer(d, to_word(act, g), out.ctypes.data_as(C.c_void_p))
        bits = np.array([(int(out[i >> 6]) >> (i & 63)) & 1 for i in range(H * H)]).reshape(H, H)
        exp = np.zeros((H, H), int)
        exp[1::2, 1::2] = act
        assert np.array_equal(bits, exp)


def test_philox_extract_select(hb):
    rng = np.random.default_rng(5)
    out = np.zeros(4, np.uint32)
    for _ in range(100):
        c = [int(x) for x in rng.integers(0, 2**32, size=6)]
        hb.hb_philox(*c, out.ctypes.data_as(C.c_void_p))
        assert np.array_equal(out, O.philox(*c))
    stream = rng.integers(0, 2**32, size=40, dtype=np.uint32)
    big = sum(int(w) << (32 * i) for i, w in enumerate(stream))
    for _ in range(500):
        n = int(rng.integers(1, 65))
        off = int(rng.integers(0, 32 * 38 - n))
        assert hb.hb_extract_bits(stream.ctypes.data_as(C.c_void_p), off, n) == (big >> off) & ((1 << n) - 1)
    for _ in range(300):
        x = int(rng.integers(1, 2**63))
        pos = [i for i in range(64) if (x >> i) & 1]
        k = int(rng.integers(0, len(pos)))
        assert hb.hb_select64(x, k) == pos[k]


def test_survey_structural_kats(hb):
    """The structural known answers SURVEY.md section 8(c) extracted from the imported reference, asserted on BOTH the oracle
    and the device helpers: d=3 centre-qubit syndromes, the d=5 plaquette-type map, logical operators, 8-neighbourhoods."""
    # d = 3, centre qubit (1,1): X -> {(1,2),(2,1)}, Z -> {(1,1),(2,2)}, Y -> all four
    o3 = O.OracleVecEnv(3, "DP", True, 3, 0.01, 0.01, 1, 0)
    want = {1: {(1, 2), (2, 1)}, 3: {(1, 1), (2, 2)}, 2: {(1, 1), (2, 2), (1, 2), (2, 1)}}
    for pauli, cells in want.items():
        hidden = np.zeros((3, 3), np.int8); hidden[1, 1] = pauli
        syn, _ = o3.syndrome_of(hidden)
        assert {tuple(x) for x in np.argwhere(syn)} == cells
        xb = to_word((hidden == 1) | (hidden == 2), 4); zb = to_word((hidden == 3) | (hidden == 2), 4)
        got = from_word(hb.hb_true_syndrome(3, xb, zb), 4, 4, 4)
        assert {tuple(x) for x in np.argwhere(got)} == cells
    # d = 5 plaquette types (1 = flips on Z/Y, 3 = flips on X/Y, 0 = absent)
    o5 = O.OracleVecEnv(5, "DP", False, 5, 0.01, 0.01, 1, 0)
    a, b, ty = o5.stab_order()
    tmap = np.zeros((6, 6), int); tmap[a, b] = ty
    assert tmap.tolist() == [[0, 3, 0, 3, 0, 0], [0, 1, 3, 1, 3, 1], [1, 3, 1, 3, 1, 0], [0, 1, 3, 1, 3, 1], [1, 3, 1, 3, 1, 0], [0, 0, 3, 0, 3, 0]]
    t1 = from_word(hb.hb_masks(5, 1), 6, 6, 6); t3 = from_word(hb.hb_masks(5, 2), 6, 6, 6)
    assert np.array_equal(t1 * 1 + t3 * 3, tmap)
    # logical operators: a full row of X has zero syndrome and label index 1, a full column of Z label index 2
    row_x = np.zeros((5, 5), np.int8); row_x[2, :] = 1
    col_z = np.zeros((5, 5), np.int8); col_z[:, 3] = 3
    for hidden, label in ((row_x, 1), (col_z, 2)):
        syn, lab = o5.syndrome_of(hidden)
        assert not syn.any() and lab == label
        xb = to_word((hidden == 1) | (hidden == 2), 6); zb = to_word((hidden == 3) | (hidden == 2), 6)
        assert hb.hb_true_syndrome(5, xb, zb) == 0 and hb.hb_label(5, xb, zb) == label
    # qubit_neighbours[0] = [1, 5, 6]; centre qubit 12 -> [11, 13, 7, 6, 8, 17, 16, 18]
    for q, nb in ((0, {1, 5, 6}), (12, {11, 13, 7, 6, 8, 17, 16, 18})):
        board = np.zeros((5, 5), np.int8); board[q // 5, q % 5] = 1
        got = from_word(hb.hb_neighbours(5, to_word(board, 6)), 6, 5, 5)
        assert {int(r * 5 + c) for r, c in np.argwhere(got)} == nb
