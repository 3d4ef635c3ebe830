"""Referee ("static") decoders as lookup tables.

The reference consults a Keras MLP once per step, batch 1 (`static_decoder.predict`,
example_notebooks/Environments.py:144) and compares `argmax` with the true homology class
(:150).  The referee only ever sees the *true* syndrome, which is a function of the
d*d-1 stabilizer bits, so the whole decoder is a table; evaluating it once, exhaustively,
with one canonical fp64 evaluator makes `done` a pure integer function that is identical
on the CPU oracle and on the GPU (SURVEY section 7 "Referee exactness").

Table formats are those of `dq_env_set_referee_lut` (include/dq_decoding.h):
  JOINT  one table over all stabilizers in `joint_order`, 2-bit entries holding X + 2Z
  SPLIT  table A over the type-3 (X-sensitive) stabilizers -> X bit,
         table B over the type-1 (Z-sensitive) stabilizers -> Z bit

Builders:
  from_keras_mlp      exhaustive evaluation of a shipped referee (`nn_d5_X_p5`, `nn_d5_DP_p5`)
  min_weight          minimum-weight decoder by breadth-first search over syndromes, for the
                      distances the reference ships no referee for (d=3, d=7; README.md:278
                      notes any homology-class decoder may be plugged in)
  from_predict        any object with the reference's duck-typed `.predict`
"""
import os
import zlib

import numpy as np

from . import h5lite

JOINT, SPLIT = 0, 1
DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


# ---- lattice tables (host side, numpy; spec = Function_Library.py:23-61, :186-233) -----------
def plaquette_present(d, a, b):
    return not ((a == 0 and b % 2 == 0) or (a == d and b % 2 == 1) or
                (b == 0 and a % 2 == 1) or (b == d and a % 2 == 0))


def stabilizer_order(d):
    """[(a, b)] in the draw order of generate_faulty_syndrome: bulk row-major, top, bottom, left, right."""
    g, nb = d + 1, (d - 1) // 2
    order = [(a, b) for a in range(1, d) for b in range(1, d)]
    order += [(0, 2 * x + 1) for x in range(nb)] + [(d, 2 * x + 2) for x in range(nb)]
    order += [(2 * x + 2, 0) for x in range(nb)] + [(2 * x + 1, d) for x in range(nb)]
    assert all(plaquette_present(d, a, b) for a, b in order) and len(order) == d * d - 1
    assert g == d + 1
    return order


def joint_order(d):
    """Bit order of a JOINT table index: grid rows 1..d-1 left to right (d stabilizers each), then the
    top/bottom boundary stabilizers by column (odd columns sit on the top row, even ones on the bottom)."""
    order = [(a, b) for a in range(1, d) for b in range(d + 1) if plaquette_present(d, a, b)]
    order += [((0 if b % 2 else d), b) for b in range(1, d)]
    assert sorted(order) == sorted(stabilizer_order(d))
    return order


def type_order(d, odd):
    """Stabilizers of one type (odd=1: type 3, flips on X/Y) in draw order."""
    return [(a, b) for a, b in stabilizer_order(d) if (a + b) % 2 == odd]


def pack2(classes):
    """uint8 class array (values 0..3) -> 2-bit packed table."""
    c = np.asarray(classes, np.uint8)
    pad = (-len(c)) % 4
    if pad:
        c = np.concatenate([c, np.zeros(pad, np.uint8)])
    c = c.reshape(-1, 4)
    return (c[:, 0] | (c[:, 1] << 2) | (c[:, 2] << 4) | (c[:, 3] << 6)).astype(np.uint8)


def unpack2(table, n):
    t = np.asarray(table, np.uint8)
    out = np.stack([(t >> s) & 3 for s in (0, 2, 4, 6)], axis=1).reshape(-1)
    return out[:n]


class RefereeLUT:
    """Packed referee tables + the reference's `.predict` face (so the same object can be handed
    to the unmodified reference env as `static_decoder`)."""

    def __init__(self, d, error_model, mode, lut_a, lut_b=None, source=""):
        self.d, self.error_model, self.mode = d, error_model, mode
        self.lut_a = np.ascontiguousarray(lut_a, np.uint8)
        self.lut_b = None if lut_b is None else np.ascontiguousarray(lut_b, np.uint8)
        self.source = source
        self.n_classes = 2 if error_model == "X" else 4
        self._order = joint_order(d)
        self._dev = {}

    # -- host-side evaluation (used by .predict and by tests) --
    def classify(self, syndrome):
        """syndrome: (d+1, d+1) or flat (d+1)^2 array of 0/1 -> class index X + 2Z."""
        s = np.asarray(syndrome).reshape(self.d + 1, self.d + 1)
        if self.mode == JOINT:
            idx = 0
            for k, (a, b) in enumerate(self._order):
                idx |= int(s[a, b]) << k
            return int((self.lut_a[idx >> 2] >> ((idx & 3) * 2)) & 3)
        out = 0
        for odd, lut, shift in ((1, self.lut_a, 0), (0, self.lut_b, 1)):
            if lut is None or (shift == 1 and self.error_model == "X"):
                continue
            idx = 0
            for k, (a, b) in enumerate(type_order(self.d, odd)):
                idx |= int(s[a, b]) << k
            out |= int((lut[idx >> 2] >> ((idx & 3) * 2)) & 1) << shift
        return out

    def predict(self, x, batch_size=1, verbose=0):
        x = np.asarray(x)
        out = np.zeros((len(x), self.n_classes), np.float32)
        for r, vec in enumerate(x):
            out[r, self.classify(vec)] = 1.0
        return out

    # -- device residency --
    def device_tables(self, device):
        import torch
        key = str(device)
        if key not in self._dev:
            a = torch.from_numpy(self.lut_a).to(device)
            b = None if self.lut_b is None else torch.from_numpy(self.lut_b).to(device)
            self._dev[key] = (a, b)
        return self._dev[key]

    # -- persistence (zlib'd raw tables; a few hundred KB for the d=5 DP referee) --
    def save(self, path):
        hdr = "DQREF1 %d %s %d %d %d %s\n" % (self.d, self.error_model, self.mode, len(self.lut_a),
                                              0 if self.lut_b is None else len(self.lut_b), self.source.replace(" ", "_"))
        payload = self.lut_a.tobytes() + (b"" if self.lut_b is None else self.lut_b.tobytes())
        with open(path, "wb") as f:
            f.write(hdr.encode())
            f.write(zlib.compress(payload, 9))

    @classmethod
    def load(cls, path):
        with open(path, "rb") as f:
            hdr = f.readline().decode().split()
            payload = zlib.decompress(f.read())
        if hdr[0] != "DQREF1":
            raise ValueError("not a referee table file: %s" % path)
        d, model, mode, na, nb = int(hdr[1]), hdr[2], int(hdr[3]), int(hdr[4]), int(hdr[5])
        a = np.frombuffer(payload, np.uint8, na).copy()
        b = np.frombuffer(payload, np.uint8, nb, na).copy() if nb else None
        return cls(d, model, mode, a, b, source=hdr[6] if len(hdr) > 6 else "")


# ---- builders --------------------------------------------------------------------------------
def load_keras_mlp(path):
    """[(kernel(in,out), bias)] of a Keras Sequential Dense stack saved with model.save()."""
    f = h5lite.H5File(path)
    names = [n for n in f.keys("/model_weights") if n.startswith("dense") and f.keys("/model_weights/" + n)]
    names.sort(key=lambda s: int(s.split("_")[-1]))
    return [(f["/model_weights/%s/%s/kernel:0" % (n, n)], f["/model_weights/%s/%s/bias:0" % (n, n)]) for n in names]


def _syndrome_vectors(d, positions, lo, hi, dtype):
    """Rows lo..hi-1 of the exhaustive input matrix: bit k of the row index sets grid cell positions[k]."""
    idx = np.arange(lo, hi, dtype=np.int64)
    x = np.zeros((hi - lo, (d + 1) * (d + 1)), dtype)
    for k, (a, b) in enumerate(positions):
        x[:, a * (d + 1) + b] = (idx >> k) & 1
    return x


def from_keras_mlp(path, d, error_model, device="cpu", batch=1 << 16, progress=None):
    """Exhaustive argmax of the shipped referee (ReLU MLP, softmax head; dropout inactive at predict).
    Canonical evaluator: float64 matmuls in the given order, argmax of the logits, ties -> lowest class."""
    import torch
    layers = [(torch.from_numpy(k.astype(np.float64)).to(device), torch.from_numpy(b.astype(np.float64)).to(device))
              for k, b in load_keras_mlp(path)]
    if error_model == "X":
        positions, mode = type_order(d, 1), SPLIT       # only X-sensitive stabilizers can fire
    else:
        positions, mode = joint_order(d), JOINT
    n = 1 << len(positions)
    classes = np.empty(n, np.uint8)
    with torch.no_grad():
        for lo in range(0, n, batch):
            hi = min(n, lo + batch)
            h = torch.from_numpy(_syndrome_vectors(d, positions, lo, hi, np.float64)).to(device)
            for i, (k, b) in enumerate(layers):
                h = h @ k + b
                if i + 1 < len(layers):
                    h = torch.relu(h)
            classes[lo:hi] = torch.argmax(h, dim=1).to("cpu").numpy().astype(np.uint8)
            if progress:
                progress(hi, n)
    return RefereeLUT(d, error_model, mode, pack2(classes), None, source="keras:" + os.path.basename(path))


def from_predict(static_decoder, d, error_model, batch=1 << 14):
    """Tabulate any object with the reference's `.predict(x[B,(d+1)^2]) -> [B,n_classes]`."""
    if error_model == "X":
        positions, mode = type_order(d, 1), SPLIT
    else:
        positions, mode = joint_order(d), JOINT
        if len(positions) > 26:
            raise ValueError("joint tabulation needs d*d-1 <= 26 stabilizers")
    n = 1 << len(positions)
    classes = np.empty(n, np.uint8)
    for lo in range(0, n, batch):
        hi = min(n, lo + batch)
        out = static_decoder.predict(_syndrome_vectors(d, positions, lo, hi, np.float32), batch_size=hi - lo, verbose=0)
        classes[lo:hi] = np.argmax(np.asarray(out), axis=1)
    return RefereeLUT(d, error_model, mode, pack2(classes), None, source="predict")


def _min_weight_bits(d, odd):
    """Class bit of a minimum-weight error for every syndrome of one stabilizer type.

    odd=1: X errors seen by type-3 stabilizers, class bit = X parity on column 0;
    odd=0: Z errors seen by type-1 stabilizers, class bit = Z parity on row 0
    (example_notebooks/Function_Library.py:322-334).  Breadth-first search by error weight from
    the trivial syndrome; among equal-weight errors the first found in (qubit-ascending) order wins.
    """
    pos = {ab: k for k, ab in enumerate(type_order(d, odd))}
    flips, cls = [], []
    for r in range(d):
        for c in range(d):
            m = 0
            for a, b in ((r, c), (r, c + 1), (r + 1, c), (r + 1, c + 1)):
                if (a, b) in pos:
                    m |= 1 << pos[(a, b)]
            flips.append(m)
            cls.append(int(c == 0) if odd else int(r == 0))
    n = 1 << len(pos)
    out = np.zeros(n, np.uint8)
    seen = np.zeros(n, bool)
    seen[0] = True
    frontier = np.array([0], np.int64)
    while len(frontier):
        cand_s, cand_c = [], []
        fc = out[frontier]
        for m, cb in zip(flips, cls):
            cand_s.append(frontier ^ m)
            cand_c.append(fc ^ cb)
        s = np.concatenate(cand_s)
        c = np.concatenate(cand_c)
        keep = ~seen[s]
        s, c = s[keep], c[keep]
        s, first = np.unique(s, return_index=True)
        out[s] = c[first]
        seen[s] = True
        frontier = s
    assert seen.all()
    return out


def min_weight(d, error_model):
    a = pack2(_min_weight_bits(d, 1))
    b = pack2(_min_weight_bits(d, 0)) if error_model == "DP" else None
    return RefereeLUT(d, error_model, SPLIT, a, b, source="min_weight")


def shipped(d, error_model):
    """The referee the package ships for (d, error_model): the tabulated reference MLP for d=5,
    the minimum-weight decoder otherwise (cached under data/)."""
    os.makedirs(DATA_DIR, exist_ok=True)
    path = os.path.join(DATA_DIR, "referee_d%d_%s.lut" % (d, error_model))
    if os.path.exists(path):
        return RefereeLUT.load(path)
    if d == 5:
        raise FileNotFoundError(path + " (tabulated from the reference's referee by tools/build_referee_luts.py)")
    ref = min_weight(d, error_model)
    ref.save(path)
    return ref
