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

for tag, sub in (("dp", "d5_dp"), ("x", "d5_x")):
    w = Q.load_keras_dqn_weights("/root/reference/trained_models/%s/0.007/final_dqn_weights.h5f" % sub)
    arrays = {}
    for i, (k, b) in enumerate(w["conv"]):
        arrays["conv%d_k" % i], arrays["conv%d_b" % i] = k, b
    for i, (k, b) in enumerate(w["dense"]):
        arrays["dense%d_k" % i], arrays["dense%d_b" % i] = k, b
    out = os.path.join(HERE, "dqn_d5_%s_0.007.npz" % tag)
    np.savez_compressed(out, **arrays)
    print(out, os.path.getsize(out))
