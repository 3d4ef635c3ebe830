"""Generates tests/golden/history_pins.npz and tests/golden/referee_pins.npz from files the reference ships.

Run in the build container (needs /root/reference):   python tests/golden/make_golden_pins.py

history_pins.npz -- for each of the 14 shipped agents (trained_models/d5_{x,dp}/<p>/training_history.json + the two config
  pickles): the per-episode lifetimes (not logged by the fork; recovered EXACTLY from consecutive `episode_lifetimes_rolling_avg`
  values -- they come out integral and multiples of volume_depth), `nb_steps`, and the columns the fork derived from them
  (`episode_lifetimes_rolling_avg`, `best_rolling_avg`, `best_episode`, `time_since_best`, `has_succeeded`, `stopped_improving`),
  plus the first non-NaN `mean_eps` and the hyper-parameters fit() was called with (Single_Point_Training_Script.py:138-152).
  tests/test_history_pins.py replays the lifetimes through deepq_decoding_b200.episodes.EpisodeBook and LinearAnnealedPolicy.
referee_pins.npz -- true syndromes and the class the shipped Keras referee MLPs (example_notebooks/referee_decoders/nn_d5_*_p5,
  evaluated in float64 from the HDF5 weights) assign to them: all 2^12 type-3 syndromes for the X referee, 200 000 random
  24-bit syndromes for the depolarising one.  tests/test_referee_pins.py checks the shipped lookup tables
  (deepq_decoding_b200/data/referee_d5_*.lut) against them; with /root/reference present it re-evaluates the MLPs as well.
"""
import glob
import json
import os
import pickle
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def recover_lifetimes(rolling, L):
    life = np.zeros(len(rolling))
    for k in range(len(rolling)):
        tot = rolling[k] * min(k + 1, L)
        prev = rolling[k - 1] * min(k, L) if k else 0.0
        life[k] = tot - prev + (life[k - L] if k >= L else 0.0)
    li = np.rint(life)
    assert np.abs(life - li).max() < 1e-6, "rolling averages do not decompose into integral lifetimes"
    return li.astype(np.int32)


def history_pins():
    out = {}
    names = []
    for f in sorted(glob.glob(os.path.join(REF, "trained_models/*/*/training_history.json"))):
        folder = os.path.dirname(f)
        model, rate = folder.split("/")[-2:]
        fixed = pickle.load(open(os.path.join(os.path.dirname(folder), "fixed_config.p"), "rb"))
        var = pickle.load(open(glob.glob(os.path.join(folder, "variable_config_*.p"))[0], "rb"))
        h = json.load(open(f))
        L = fixed["rolling_average_length"]
        name = "%s_%s" % (model, rate)
        names.append(name)
        life = recover_lifetimes(np.array(h["episode_lifetimes_rolling_avg"]), L)
        assert (life % fixed["volume_depth"] == 0).all() and life.min() > 0
        first = next(i for i, x in enumerate(h["mean_eps"]) if x == x)
        out[name + "/lifetime"] = life
        out[name + "/nb_steps"] = np.array(h["nb_steps"], np.int64)
        out[name + "/rolling"] = np.array(h["episode_lifetimes_rolling_avg"], np.float64)
        out[name + "/best_rolling"] = np.array(h["best_rolling_avg"], np.float64)
        out[name + "/best_episode"] = np.array(h["best_episode"], np.int32)
        out[name + "/time_since_best"] = np.array(h["time_since_best"], np.int32)
        out[name + "/has_succeeded"] = np.array(h["has_succeeded"], bool)
        out[name + "/stopped_improving"] = np.array(h["stopped_improving"], bool)
        out[name + "/first_eps"] = np.array([first, h["nb_steps"][first], h["nb_steps"][first - 1] if first else 0, h["mean_eps"][first]], np.float64)
        out[name + "/hyper"] = np.array([L, var["success_threshold"], fixed["stopping_patience"], var["exploration_fraction"],
                                         var["max_eps"], var["final_eps"], var["learning_starts"], fixed["max_timesteps"]], np.float64)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "history_pins.npz"), **out)
    print("history_pins.npz: %d agents, %d bytes" % (len(names), os.path.getsize(os.path.join(HERE, "history_pins.npz"))))


def mlp_classes(path, x):
    """argmax of the shipped Keras MLP (Dense+ReLU ... Dense+softmax; Dropout inactive at predict), numpy float64."""
    from deepq_decoding_b200 import referee as R
    layers = [(np.asarray(W, np.float64), np.asarray(b, np.float64)) for W, b in R.load_keras_mlp(path)]
    h = x.astype(np.float64)
    for i, (W, b) in enumerate(layers):
        h = h @ W + b
        if i + 1 < len(layers):
            h = np.maximum(h, 0.0)
    return np.argmax(h, axis=1).astype(np.uint8)


def referee_pins():
    from deepq_decoding_b200 import referee as R
    d, g = 5, 6
    out = {}
    rng = np.random.default_rng(20261017)
    for model, fn in (("X", "nn_d5_X_p5"), ("DP", "nn_d5_DP_p5")):
        path = os.path.join(REF, "example_notebooks/referee_decoders", fn)
        if model == "X":
            order = R.type_order(d, 1)                       # bit-flip noise only ever lights type-3 stabilizers
            idx = np.arange(1 << len(order), dtype=np.int64)
        else:
            order = R.stabilizer_order(d)
            idx = rng.integers(0, 1 << len(order), size=200000, dtype=np.int64)
        x = np.zeros((len(idx), g * g), np.float64)
        for k, (a, b) in enumerate(order):
            x[:, a * g + b] = (idx >> k) & 1
        cls = np.concatenate([mlp_classes(path, x[i:i + 8192]) for i in range(0, len(x), 8192)])
        out[model + "/order"] = np.array(order, np.int32)
        out[model + "/index"] = idx.astype(np.uint32)
        out[model + "/class"] = cls
        print(model, "classes", np.bincount(cls))
    np.savez_compressed(os.path.join(HERE, "referee_pins.npz"), **out)
    print("referee_pins.npz: %d bytes" % os.path.getsize(os.path.join(HERE, "referee_pins.npz")))


if __name__ == "__main__":
    history_pins()
    referee_pins()
