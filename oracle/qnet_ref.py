"""fp32 torch restatement of the reference Q-network and DQN update (TEST INFRASTRUCTURE).

Reference followed:
  network ......... example_notebooks/Function_Library.py:338-377 (build_convolutional_nn): Conv2D stack
                    (valid padding, channels_first, ReLU), Flatten (C,H,W order), Dense+ReLU(+Dropout) stack,
                    Dense(num_actions) linear; keras-rl then stacks the dueling head Dense(num_actions+1) and
                    Q = a0 + a[1:] - mean(a[1:])   (dueling_type='avg'; SURVEY 8a row 15, shapes read from
                    trained_models/d5_dp/0.007/final_dqn_weights.h5f)
  weights ......... Keras HDF5: conv kernels HWIO, Dense (in,out)
  DQN update ...... keras-rl 0.4.x agents/dqn.py semantics (the fork's source is not in /root/reference;
                    SURVEY 8a row 19): double-DQN target, 0.5*(y-Q[a])^2 mean over the batch, Keras Adam
                    (eps inside the sqrt's sum: p -= lr_t*m/(sqrt(v)+eps), lr_t = lr*sqrt(1-b2^t)/(1-b1^t))
  policies ........ SURVEY 8a rows 16-17

Parity is pinned by the notebook-3 greedy-decode KAT and by the published lifetimes (tests/test_qnet_*),
not bit-for-bit ("parity unpinned" at the bit level for the keras-rl fork: its source is unavailable).
"""
import numpy as np
import torch
import torch.nn.functional as F


def load_keras_dqn_weights(path):
    """final_dqn_weights.h5f -> dict of numpy arrays in Keras layouts."""
    from deepq_decoding_b200.h5lite import H5File
    f = H5File(path)
    out = {"conv": [], "dense": []}
    for name in sorted(k for k in f.keys("/") if k.startswith("conv2d_") and f.keys("/" + k)):
        g = "/%s/%s" % (name, name)
        out["conv"].append((f[g + "/kernel:0"], f[g + "/bias:0"]))
    dn = sorted((k for k in f.keys("/") if k.startswith("dense_") and f.keys("/" + k)), key=lambda s: int(s.split("_")[1]))
    for name in dn:
        sub = f.keys("/" + name)[0]
        g = "/%s/%s" % (name, sub)
        out["dense"].append((f[g + "/kernel:0"], f[g + "/bias:0"]))
    return out


class TorchQNet:
    """Plain fp32 forward/backward of the reference architecture from Keras-layout weights."""

    def __init__(self, conv, dense, strides, dueling=True):
        # conv: [(HWIO kernel, bias)], dense: [(in,out kernel, bias)] (the last one is the dueling head when dueling)
        self.conv = [(torch.tensor(k).permute(3, 2, 0, 1).contiguous().float().requires_grad_(True),
                      torch.tensor(b).float().requires_grad_(True)) for k, b in conv]
        self.dense = [(torch.tensor(k).float().requires_grad_(True), torch.tensor(b).float().requires_grad_(True))
                      for k, b in dense]
        self.strides, self.dueling = list(strides), dueling

    def parameters(self):
        return [t for pair in self.conv + self.dense for t in pair]

    def forward(self, obs, dropout_masks=None):
        """obs [B,C,H,W] (0/1) -> Q [B,A].  dropout_masks: optional list (per hidden dense layer) of
        [B,units] multiplicative masks (already scaled by 1/keep) -- training mode."""
        x = torch.as_tensor(obs).float()
        for (k, b), s in zip(self.conv, self.strides):
            x = F.relu(F.conv2d(x, k, b, stride=s))
        x = x.flatten(1)                       # C,H,W order (Keras channels_first Flatten)
        nd = len(self.dense) - (2 if self.dueling else 1)
        for i, (k, b) in enumerate(self.dense):
            x = x @ k + b
            if i < nd:
                x = F.relu(x)
                if dropout_masks is not None and dropout_masks[i] is not None:
                    x = x * dropout_masks[i]
        if self.dueling:
            x = x[:, :1] + x[:, 1:] - x[:, 1:].mean(dim=1, keepdim=True)
        return x


def glorot_uniform_params(rng, in_channels, conv_cfg, dense_units, num_actions, in_side, dueling=True):
    """Keras default init (glorot_uniform kernels, zero biases) in Keras layouts, from a numpy Generator."""
    conv, c, side = [], in_channels, in_side
    for filters, ksz, stride in conv_cfg:
        fan_in, fan_out = c * ksz * ksz, filters * ksz * ksz
        lim = np.sqrt(6.0 / (fan_in + fan_out))
        conv.append((rng.uniform(-lim, lim, size=(ksz, ksz, c, filters)).astype(np.float32), np.zeros(filters, np.float32)))
        c, side = filters, (side - ksz) // stride + 1
    dense, n_in = [], c * side * side
    outs = list(dense_units) + [num_actions] + ([num_actions + 1] if dueling else [])
    for units in outs:
        lim = np.sqrt(6.0 / (n_in + units))
        dense.append((rng.uniform(-lim, lim, size=(n_in, units)).astype(np.float32), np.zeros(units, np.float32)))
        n_in = units
    return conv, dense


def keras_adam_step(params, grads, m, v, t, lr, b1=0.9, b2=0.999, eps=1e-7):
    """In-place Keras-2 Adam on lists of tensors; t is the 1-based update index."""
    lr_t = lr * np.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t)
    with torch.no_grad():
        for p, g, mi, vi in zip(params, grads, m, v):
            mi.mul_(b1).add_(g, alpha=1 - b1)
            vi.mul_(b2).addcmul_(g, g, value=1 - b2)
            p.sub_(lr_t * mi / (vi.sqrt() + eps))


def dqn_targets(q_online_next, q_target_next, reward, terminal, gamma):
    """Double-DQN: y = r + gamma*(1-terminal)*Q_target(s', argmax_a Q_online(s', a))."""
    a_star = q_online_next.argmax(dim=1)
    boot = q_target_next.gather(1, a_star[:, None])[:, 0]
    return reward + gamma * (1.0 - terminal) * boot


def dqn_loss(q, actions, y):
    """0.5*(y - Q[a])^2, mean over the batch (delta_clip = inf)."""
    qa = q.gather(1, actions[:, None].long())[:, 0]
    return 0.5 * ((y - qa) ** 2).mean()


def masked_argmax(q, legal_mask_words, num_actions):
    """argmax of Q restricted to the legal actions (GreedyQPolicy(masked_greedy=True)); ties -> lowest index."""
    q = np.asarray(q, np.float32)
    out = np.zeros(len(q), np.int32)
    for i in range(len(q)):
        best, arg = -np.inf, num_actions - 1
        for a in range(num_actions):
            if (int(legal_mask_words[i][a >> 6]) >> (a & 63)) & 1 and q[i, a] > best:
                best, arg = q[i, a], a
        out[i] = arg
    return out
