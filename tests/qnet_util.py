import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def golden_weights(tag):
    z = np.load(os.path.join(HERE, "golden", "dqn_d5_%s_0.007.npz" % tag))
    conv = [(z["conv%d_k" % i], z["conv%d_b" % i]) for i in range(3)]
    dense = [(z["dense%d_k" % i], z["dense%d_b" % i]) for i in range(3)]
    return conv, dense


REF_CC = [[64, 3, 2], [32, 2, 1], [32, 2, 1]]          # cluster_scripts/d5_dp/0.001/Single_Point_Training_Script.py:61-90
REF_FF = [[512, 0.2]]
