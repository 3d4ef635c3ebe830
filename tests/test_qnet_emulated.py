"""CPU check of the fp32 SIMT half of dq_qnet.cu -- the kernel SOURCE compiled by g++ under tests/host/cuda_emu.h (launches
and dynamic shared memory rewritten by tests/host/emu_build.py, the tcgen05 path cut off) -- against the torch restatement
(oracle/qnet_ref.py).  Mirrors tests/test_qnet_gpu.py at sizes the fibers finish in seconds; same tolerances.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import oracle as O
from oracle import qnet_ref as QR
from qnet_util import REF_CC, golden_weights, Bf16SimQNet as _Bf16SimQNet, dropout_mask as _dropout_mask
import emu_qnet as EQ

Q_ATOL, G_RTOL = 2e-3, 2e-3
p = EQ._p


def random_boards(n, channels, seed, side=11, density=0.12):
    rng = np.random.default_rng(seed)
    return (rng.random((n, channels, side, side)) < density).astype(np.uint8)


def test_emulated_forward_matches_torch_with_shipped_agent():
    conv, dense = golden_weights("dp")
    net = QR.TorchQNet(conv, dense, strides=[2, 1, 1])
    q = EQ.EmuQNet(REF_CC, [[512, 0.2]], (7, 11, 11), 51, max_batch=96)
    q.set_keras_weights(conv, dense)
    boards = random_boards(96, 7, 1)
    got, packed = q.forward(boards)
    want = net.forward(boards).detach().numpy()
    assert np.abs(got - want).max() < Q_ATOL
    assert (got.argmax(1) == want.argmax(1)).mean() > 0.98
    # the pack kernel: bit x*H+y of layer l of sample b in packed[l*PW + word][b]
    from deepq_decoding_b200.envs import unpack_observations
    assert np.array_equal(unpack_observations(packed, 96, 5, 7), boards)


@pytest.mark.parametrize("cfg", [(REF_CC, [(512, 0.0)], 7, 51, 48), ([[16, 3, 2], [8, 2, 1]], [(32, 0.0), (24, 0.0)], 6, 26, 40)],
                         ids=["reference_net", "small_two_dense"])
def test_emulated_backward_matches_autograd(cfg):
    cc, ff, channels, A, B = cfg
    rng = np.random.default_rng(7)
    conv, dense = QR.glorot_uniform_params(rng, channels, cc, [u for u, _ in ff], A, 11)
    for _, b in conv + dense:                       # non-zero biases so that their gradients are exercised too
        b += rng.standard_normal(b.shape).astype(np.float32) * 0.05
    q = EQ.EmuQNet(cc, [[u, 0.0] for u, _ in ff], (channels, 11, 11), A, max_batch=B)
    q.set_keras_weights(conv, dense)
    net = QR.TorchQNet(conv, dense, strides=[l[2] for l in cc])
    boards = random_boards(B, channels, 2, density=0.2)
    dq = rng.standard_normal((B, A)).astype(np.float32)
    got_q, packed = q.forward(boards, train=True)
    want_q = net.forward(boards)
    assert np.abs(got_q - want_q.detach().numpy()).max() < 1e-3
    (want_q * torch.tensor(dq)).sum().backward()
    got = q.keras_grads(q.backward(packed, dq))
    for t, (g, prm) in enumerate(zip(got, net.parameters())):
        want = prm.grad.numpy()
        assert g.shape == want.shape
        assert np.abs(g - want).max() <= G_RTOL * max(1e-6, np.abs(want).max()), "gradient of tensor %d" % t


def test_emulated_dqn_update_matches_oracle():
    """Double-DQN target, 0.5*err^2 mean over the batch, Keras Adam: parameters after 3 updates."""
    L = EQ.lib()
    B, nA, gamma, lr = 32, 51, 0.99, 1e-3
    rng = np.random.default_rng(5)
    conv, dense = QR.glorot_uniform_params(rng, 7, REF_CC, [512], nA, 11)
    q = EQ.EmuQNet(REF_CC, [[512, 0.0]], (7, 11, 11), nA, max_batch=B)
    q.set_keras_weights(conv, dense)
    online = QR.TorchQNet(conv, dense, strides=[2, 1, 1])
    tconv, tdense = [(k * 0.9, b * 0.9) for k, b in conv], [(k * 0.9, b * 0.9) for k, b in dense]
    target = QR.TorchQNet(tconv, tdense, strides=[2, 1, 1])
    online_params = q.params
    qt = EQ.EmuQNet(REF_CC, [[512, 0.0]], (7, 11, 11), nA, max_batch=B)
    qt.set_keras_weights(tconv, tdense)
    m, v = np.zeros_like(q.params), np.zeros_like(q.params)
    tm = [torch.zeros_like(x) for x in online.parameters()]
    tv = [torch.zeros_like(x) for x in online.parameters()]
    for t in range(1, 4):
        s0, s1 = random_boards(B, 7, 10 + t), random_boards(B, 7, 20 + t)
        act = rng.integers(0, nA, size=B).astype(np.int32)
        rew = (rng.random(B) < 0.3).astype(np.float32)
        term = (rng.random(B) < 0.2).astype(np.uint8)
        qo_next, _ = q.forward(s1)
        qt_next, _ = qt.forward(s1)
        y = np.zeros(B, np.float32)
        EQ.check(L.dq_dqn_targets(p(qo_next), p(qt_next), p(rew), p(term), gamma, B, nA, p(y), None))
        q0, packed0 = q.forward(s0, train=True)
        dq, stats = np.zeros((B, nA), np.float32), np.zeros(2, np.float32)
        EQ.check(L.dq_dqn_loss_grad(p(q0), p(act), p(y), B, nA, p(dq), p(stats), None))
        grads = q.backward(packed0, dq)
        EQ.check(L.dq_adam_step(p(q.params), p(m), p(v), p(grads), q.num_params, lr, 0.9, 0.999, 1e-7, t, 1.0, None))
        with torch.no_grad():
            ty = QR.dqn_targets(online.forward(s1), target.forward(s1), torch.tensor(rew), torch.tensor(term).float(), gamma)
        assert np.abs(y - ty.numpy()).max() < 1e-4
        for prm in online.parameters():
            prm.grad = None
        loss = QR.dqn_loss(online.forward(s0), torch.tensor(act), ty)
        loss.backward()
        assert abs(float(stats[0]) / B - float(loss.detach())) < 1e-4 * max(1.0, abs(float(loss.detach())))     # stats[0] = sum of per-sample losses
        QR.keras_adam_step(online.parameters(), [prm.grad for prm in online.parameters()], tm, tv, t, lr)
    assert online_params is q.params
    got = q.keras_grads(q.params)                 # same tensor order / layouts as TorchQNet.parameters()
    for g, prm in zip(got, online.parameters()):
        assert np.abs(g - prm.detach().numpy()).max() < 2e-5


def test_emulated_eps_greedy_policy_exact():
    L = EQ.lib()
    rng = np.random.default_rng(0)
    n, A, seed, base, step = 300, 51, 0x1234567812345678, 40, 9
    q = rng.standard_normal((n, A)).astype(np.float32)
    q[::7, 3] = q[::7, 11] = 9.0                                    # ties -> lowest index
    legal = rng.integers(0, 2**51, size=(n, 1), dtype=np.uint64) | np.uint64(1 << 50)
    out = np.zeros(n, np.int32)
    for eps, masked in ((0.0, 1), (0.0, 0), (0.3, 0), (1.0, 1)):
        EQ.check(L.dq_policy_eps_greedy(p(q), p(legal), n, 1, A, base, seed, step, None, eps, masked, p(out), None))
        thr = 0 if eps <= 0 else min(int(np.floor(eps * 2**32)), 2**32 - 1)
        for i in range(0, n, 3):
            u = O.philox(base + i, step, 0, 1, seed & 0xFFFFFFFF, seed >> 32)
            bits = [a for a in range(A) if (int(legal[i, 0]) >> a) & 1]
            if int(u[1]) < thr:
                want = bits[(int(u[0]) * len(bits)) >> 32]
            else:
                cand = bits if masked else range(A)
                want = max(cand, key=lambda a: (q[i, a], -a))
            assert out[i] == want, (eps, masked, i)


def test_emulated_replay_sample_gathers_consecutive_slots():
    L = EQ.lib()
    rows, npad, n, cap, B = 14, 64, 50, 6, 256
    rng = np.random.default_rng(1)
    obs = rng.integers(0, 2**62, size=(cap, rows, npad)).astype(np.uint64)
    act = rng.integers(0, 51, size=(cap, n)).astype(np.int32)
    rew = rng.random((cap, n)).astype(np.float32)
    term = (rng.random((cap, n)) < 0.1).astype(np.uint8)
    head, filled = 3, cap - 1
    s0, s1 = np.zeros((rows, B), np.uint64), np.zeros((rows, B), np.uint64)
    a, r, t = np.zeros(B, np.int32), np.zeros(B, np.float32), np.zeros(B, np.uint8)
    picked = np.zeros((B, 2), np.int32)
    EQ.check(L.dq_replay_sample(p(obs), p(act), p(rew), p(term), rows, npad, n, cap, head, filled, B, 5, 0,
                                p(s0), p(s1), p(a), p(r), p(t), p(picked), None))
    assert set(picked[:, 0]) <= set(range(cap)) - {head} and picked[:, 1].max() < n and picked[:, 1].min() >= 0
    assert len(set(picked[:, 0])) == cap - 1                              # every complete slot is reachable
    for b in range(B):
        ts, i = picked[b]
        assert np.array_equal(s0[:, b], obs[ts, :, i]) and np.array_equal(s1[:, b], obs[(ts + 1) % cap, :, i])
        assert a[b] == act[ts, i] and r[b] == rew[ts, i] and t[b] == term[ts, i]


@pytest.mark.parametrize("which,A", [("dp", 51), ("x", 26)])
def test_emulated_folded_head_is_the_networks_head(which, A):
    """dq_qnet_fold_head: Dense(A) -> dueling Dense(A+1) -> 'avg' combine as one affine map, on the reference's published
    weights: h @ w + b reproduces the head applied to a hidden activation h (float64 restatement)."""
    conv, dense = golden_weights(which)
    C_in = 7 if which == "dp" else 6
    q = EQ.EmuQNet(REF_CC, [[512, 0.2]], (C_in, 11, 11), A, max_batch=64)
    q.set_keras_weights(conv, dense)
    w, b = q.fold_head()
    (w2, b2), (w3, b3) = dense[-2], dense[-1]
    rng = np.random.default_rng(3)
    h = np.maximum(rng.standard_normal((64, w2.shape[0])), 0).astype(np.float64)
    y = (h @ w2.astype(np.float64) + b2) @ w3.astype(np.float64) + b3
    want = y[:, :1] + y[:, 1:] - y[:, 1:].mean(axis=1, keepdims=True)
    got = h @ w.astype(np.float64) + b
    assert np.abs(got - want).max() < 1e-4 * max(1.0, np.abs(want).max())


# ---- bf16 training path (dq_qnet_forward_tc_train / dq_qnet_backward_tc) ----------------------------------------------------
# The tcgen05 GEMM kernels are swapped for plain loops (tests/host/tc_emu.h); everything around them runs as written.  The
# reference is the torch network with the same roundings: bf16 weights and activations, fp32 accumulation, straight-through
# gradients to the fp32 master weights -- so both sides take the same ReLU branches and the comparison is tight (the plain fp32
# network differs from ANY bf16 forward by a few percent of the gradient norm: pre-activations within rounding distance of zero
# flip their ReLU branch).
@pytest.mark.parametrize("cfg", [(REF_CC, [(512, 0.2)], 7, 51, 48, 11), ([[16, 3, 2], [8, 2, 1]], [(32, 0.0), (24, 0.25)], 6, 26, 40, 11),
                                 ([[8, 3, 2], [8, 2, 1]], [(16, 0.0)], 6, 26, 5, 11),
                                 ([[16, 3, 2], [8, 2, 1], [8, 2, 1]], [(64, 0.2)], 9, 99, 12, 15)],      # BASELINE config C5's geometry: d = 7, 9 x 15 x 15, 99 actions
                         ids=["reference_net", "small_two_dense", "ragged_batch", "d7_geometry"])
def test_emulated_bf16_training_path_matches_rounded_autograd(cfg):
    cc, ff, channels, A, B, side = cfg
    rng = np.random.default_rng(11)
    conv, dense = QR.glorot_uniform_params(rng, channels, cc, [u for u, _ in ff], A, side)
    for _, b in conv + dense:
        b += rng.standard_normal(b.shape).astype(np.float32) * 0.05
    q = EQ.EmuQNet(cc, [[u, r] for u, r in ff], (channels, side, side), A, max_batch=B, tc=True)
    q.set_keras_weights(conv, dense)
    net = _Bf16SimQNet(conv, dense, strides=[l[2] for l in cc])
    boards = random_boards(B, channels, 3, side=side, density=0.2)
    dq = rng.standard_normal((B, A)).astype(np.float32)
    packed = q.pack(boards)
    seed = 0x1234500077
    masks = [torch.tensor(_dropout_mask(B, u, r, seed, i)) if r > 0 else None for i, (u, r) in enumerate(ff)]
    # inference: folded head, no dropout
    want_inf = net.forward(boards).detach().numpy()
    assert np.abs(q.forward_tc(packed) - want_inf).max() < 2e-3
    got_q = q.forward_tc(packed, train=True, dropout_seed=seed)
    want_q = net.forward(boards, dropout_masks=masks)
    assert np.abs(got_q - want_q.detach().numpy()).max() < 1e-4
    (want_q * torch.tensor(dq)).sum().backward()
    got = q.keras_grads(q.backward_tc(packed, dq))
    for t, (g, prm) in enumerate(zip(got, net.parameters())):
        want = prm.grad.numpy()
        assert g.shape == want.shape
        # dY and the column gradients are rounded to bf16 on their way into the GEMMs: a few 1e-3 of the tensor's norm
        # (a 5-sample batch averages fewer roundings per entry: looser)
        tol = 1e-2 if B >= 40 else 3e-2
        assert np.linalg.norm(g - want) <= tol * np.linalg.norm(want) + 1e-6, (t, np.linalg.norm(g - want) / np.linalg.norm(want))


def test_emulated_bf16_backward_scratch_regrows_with_the_batch():
    """dq_qnet_backward_tc sizes its scratch for the first batch it sees and regrows it for a larger one: the gradient of a batch
    does not depend on what the handle computed before."""
    cc, ff, channels, A = [[8, 3, 2], [8, 2, 1]], [[16, 0.0]], 6, 26
    rng = np.random.default_rng(5)
    conv, dense = QR.glorot_uniform_params(rng, channels, cc, [16], A, 11)
    boards = random_boards(24, channels, 4, density=0.2)
    dq = rng.standard_normal((24, A)).astype(np.float32)

    def grads(warm_up_batch):
        q = EQ.EmuQNet(cc, ff, (channels, 11, 11), A, max_batch=24, tc=True)
        q.set_keras_weights(conv, dense)
        if warm_up_batch:
            pk = q.pack(boards[:warm_up_batch])
            q.forward_tc(pk, train=True)
            q.backward_tc(pk, dq[:warm_up_batch])
        pk = q.pack(boards)
        q.forward_tc(pk, train=True)
        return q.backward_tc(pk, dq)
    fresh, regrown = grads(0), grads(7)
    assert np.array_equal(fresh, regrown) and np.abs(fresh).max() > 0
