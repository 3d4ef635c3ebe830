"""Carries the reference's trained agents to the GPU box: Keras-layout weights of
trained_models/d5_{dp,x}/0.007/final_dqn_weights.h5f as .npz (the box has no /root/reference),
plus the notebook-3 production-decoding volume (README.md:719-829).
Run in the build container:  python tests/golden/make_golden_qnet.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import qnet_ref as Q  # noqa

# d5_dp/0.005: the agent the reference's controller hands to the 0.007 runs as initial weights (Controller.py:187-270)
for tag, sub, rate in (("dp", "d5_dp", "0.007"), ("x", "d5_x", "0.007"), ("dp", "d5_dp", "0.005")):
    w = Q.load_keras_dqn_weights("/root/reference/trained_models/%s/%s/final_dqn_weights.h5f" % (sub, rate))
    arrays = {}
    for i, (k, b) in enumerate(w["conv"]):
        arrays["conv%d_k" % i], arrays["conv%d_b" % i] = k, b
    for i, (k, b) in enumerate(w["dense"]):
        arrays["dense%d_k" % i], arrays["dense%d_b" % i] = k, b
    out = os.path.join(HERE, "dqn_d5_%s_%s.npz" % (tag, rate))
    np.savez_compressed(out, **arrays)
    print(out, os.path.getsize(out))
