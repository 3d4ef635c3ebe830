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


# ---- the bf16 training path's reference: the torch network with the SAME roundings (bf16 weights and activations, fp32
# accumulation, straight-through gradients to the fp32 master weights), so both sides take the same ReLU branches
from oracle import oracle as O          # noqa: E402
from oracle import qnet_ref as QR       # noqa: E402


def _bf16(x):
    return x + (x.bfloat16().float() - x).detach()


class Bf16SimQNet(QR.TorchQNet):
    def forward(self, obs, dropout_masks=None):
        import torch
        import torch.nn.functional as F
        x = torch.as_tensor(obs).float()
        for (k, b), s in zip(self.conv, self.strides):
            x = _bf16(F.relu(F.conv2d(x, _bf16(k), b, stride=s)))
        x = x.flatten(1)
        nd = len(self.dense) - (2 if self.dueling else 1)
        for i, (k, b) in enumerate(self.dense):
            x = x @ (_bf16(k) if i <= nd else k) + b            # the dueling layer stays fp32
            if i < nd:
                x = _bf16(F.relu(x))
                if dropout_masks is not None and dropout_masks[i] is not None:
                    x = _bf16(x * dropout_masks[i])
        if self.dueling:
            x = x[:, :1] + x[:, 1:] - x[:, 1:].mean(dim=1, keepdim=True)
        return x


def dropout_mask(B, units, rate, seed, layer):
    """dropout_kernel / dropout_bf16_kernel: element i = 4*i4 + j keeps iff word j of Philox(i4 lo, i4 hi, layer, 3; seed lo, hi) >= rate*2^32."""
    thr = min(int(np.float32(rate) * np.float32(4294967296.0)), 4294967295)
    n = B * units
    out = np.zeros(n, np.float32)
    for i4 in range((n + 3) // 4):
        w = O.philox(i4 & 0xFFFFFFFF, i4 >> 32, layer, 3, seed & 0xFFFFFFFF, seed >> 32)
        for j in range(4):
            if 4 * i4 + j < n and w[j] >= thr:
                out[4 * i4 + j] = np.float32(1.0) / (np.float32(1.0) - np.float32(rate))
    return out.reshape(B, units)


