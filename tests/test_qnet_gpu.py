"""GPU parity of the Q-network / DQN kernels (through the C ABI) against the fp32 torch restatement
(oracle/qnet_ref.py).  Floating point: tolerances are stated per test; everything integer (policy picks,
replay gathers) is exact.
"""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as O
from oracle import qnet_ref as QR
from qnet_util import REF_CC, REF_FF, golden_weights

pytestmark = pytest.mark.gpu

Q_ATOL = 2e-3            # Q values are ~30; fp32 with a different summation order
G_RTOL = 2e-3            # gradients, relative to the largest entry of the tensor


def ref_net(tag):
    conv, dense = golden_weights(tag)
    return QR.TorchQNet(conv, dense, strides=[2, 1, 1]), conv, dense


def random_boards(n, channels, seed, density=0.12):
    rng = np.random.default_rng(seed)
    return (rng.random((n, channels, 11, 11)) < density).astype(np.uint8)


def env_boards(n_steps=12, n=256):
    """Realistic inputs: observations of the CUDA env under the random-legal policy."""
    import torch
    from deepq_decoding_b200.envs import VecSurfaceCodeEnv
    env = VecSurfaceCodeEnv(5, 0.02, 0.02, "DP", False, 5, None, n_envs=n, seed=3)
    env.reset()
    out = []
    for t in range(n_steps):
        obs = env.step(env.random_legal_actions(t))[0]
        out.append(obs.clone())
    packed_rows = env.get_state_words()[7:, :n].clone()
    last = out[-1]
    env.close()
    return torch.cat(out).cpu().numpy(), last, packed_rows


def test_forward_matches_torch_with_shipped_agent():
    import torch
    from deepq_decoding_b200.qnet import QNetwork
    net, conv, dense = ref_net("dp")
    q = QNetwork(REF_CC, REF_FF, (7, 11, 11), 51, dueling=True, max_batch=4096)
    q.set_keras_weights(conv, dense)
    boards, last, packed_rows = env_boards()
    boards = np.concatenate([boards, random_boards(500, 7, 1)])
    got = q.forward(boards).cpu().numpy()
    want = net.forward(boards).detach().numpy()
    assert np.abs(got - want).max() < Q_ATOL
    assert (got.argmax(1) == want.argmax(1)).mean() > 0.999
    # intermediate activations (ours are channels-last)
    x = torch.tensor(boards[:64]).float()
    a1 = torch.relu(torch.nn.functional.conv2d(x, net.conv[0][0], net.conv[0][1], stride=2))
    q.forward(boards[:64])
    mine = q.activation(0, 64).cpu().view(64, 25, 64).permute(0, 2, 1).reshape(64, 64, 5, 5)
    assert torch.allclose(mine, a1.detach(), atol=1e-4)
    # the packed rows inside the env state feed the network directly (no byte observation)
    direct = q.forward_packed(packed_rows.data_ptr(), packed_rows.shape[1], packed_rows.shape[1]).cpu().numpy()
    assert np.abs(direct - net.forward(last.cpu().numpy()).detach().numpy()).max() < Q_ATOL
    # weights survive the Keras round trip (flatten permutation, HWIO)
    c2, d2 = q.get_keras_weights()
    assert all(np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) for a, b in zip(conv + dense, c2 + d2))
    q.close()


def test_notebook3_greedy_decode_kat(tmp_path):
    """README.md:719-829: the d5_x/0.007 agent on that volume suggests corrections == [21], then the identity."""
    from deepq_decoding_b200 import agents as A
    conv, dense = golden_weights("x")
    state = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "qnet_d5_x_kat.npz"))["state0"].astype(int)
    state[5] = 0                   # the fixture was saved after the first correction had been drawn in
    spec = A.build_convolutional_nn(REF_CC, REF_FF, (6, 11, 11), 26)
    dqn = A.DQNAgent(model=spec, nb_actions=26, memory=A.SequentialMemory(limit=100, window_length=1), nb_steps_warmup=10,
                     target_model_update=100, policy=A.GreedyQPolicy(masked_greedy=True), test_policy=A.GreedyQPolicy(masked_greedy=True),
                     gamma=0.99, enable_dueling_network=True)
    dqn.compile(A.Adam(lr=1e-3), max_envs=64)
    dqn.model.set_keras_weights(conv, dense)
    corrections, identity = [], 25
    for _ in range(5):
        a = dqn.forward(state)
        if a in corrections or a == identity:
            break
        corrections.append(a)
        state[5] = 0
        for c in corrections:
            state[5, 2 * (c // 5) + 1, 2 * (c % 5) + 1] = 1
    assert corrections == [21]
    q = dqn.model.forward(state[None]).cpu().numpy()[0]
    assert abs(q[25] - 34.91) < 0.01
    # and the HDF5 writer/reader pair keeps the agent intact
    path = str(tmp_path / "w.h5f")
    dqn.save_weights(path)
    dqn.model.init_glorot(5)
    dqn.load_weights(path)
    assert abs(dqn.model.forward(state[None]).cpu().numpy()[0][25] - 34.91) < 0.01


@pytest.mark.parametrize("cfg", [
    (REF_CC, REF_FF, 7, 51, 96),                                   # the reference architecture (DP)
    ([[16, 3, 1], [8, 2, 2]], [[40, 0.0], [24, 0.0]], 4, 13, 50),   # other kernel sizes / strides / depths
])
def test_backward_matches_autograd(cfg):
    import torch
    from deepq_decoding_b200.qnet import QNetwork
    cc, ff, channels, A, B = cfg
    q = QNetwork(cc, [[u, 0.0] for u, _ in ff], (channels, 11, 11), A, dueling=True, max_batch=128, seed=7)
    conv, dense = q.get_keras_weights()
    net = QR.TorchQNet(conv, dense, strides=[l[2] for l in cc])
    boards = random_boards(B, channels, 2, density=0.2)
    rng = np.random.default_rng(3)
    dq = rng.standard_normal((B, A)).astype(np.float32)
    got_q = q.forward(boards, train=True)
    packed = q._packed
    grads = torch.zeros(q.num_params, device="cuda")
    q.backward_packed(packed.data_ptr(), B, B, torch.tensor(dq).cuda(), grads)
    want_q = net.forward(boards)
    assert np.abs(got_q.cpu().numpy() - want_q.detach().numpy()).max() < 1e-3
    (want_q * torch.tensor(dq)).sum().backward()
    g = grads.cpu().numpy()
    perm = q._flatten_perm()
    t = 0
    for (wo, bo, K, N), (kt, bt) in zip(q.layout, net.conv + net.dense):
        if t < len(cc):
            want_w = kt.grad.permute(2, 3, 1, 0).reshape(K, N).numpy()        # OIHW -> HWIO flattened
        else:
            want_w = kt.grad.numpy()
            if t == len(cc):
                want_w = want_w[perm]
        got_w, got_b, want_b = g[wo:wo + K * N].reshape(K, N), g[bo:bo + N], bt.grad.numpy()
        assert np.abs(got_w - want_w).max() <= G_RTOL * max(1e-6, np.abs(want_w).max()), "kernel grad of tensor %d" % t
        assert np.abs(got_b - want_b).max() <= G_RTOL * max(1e-6, np.abs(want_b).max()), "bias grad of tensor %d" % t
        t += 1
    q.close()


def test_dqn_update_matches_oracle():
    """One keras-rl backward(): double-DQN target, 0.5*err^2 mean, Keras Adam -- parameters after 3 updates."""
    import torch
    from deepq_decoding_b200 import agents as A
    B, nA = 64, 51
    spec = A.build_convolutional_nn(REF_CC, [[512, 0.0]], (7, 11, 11), nA)
    dqn = A.DQNAgent(model=spec, nb_actions=nA, memory=A.SequentialMemory(limit=1000), gamma=0.99, enable_dueling_network=True,
                     batch_size=B, seed=11)
    dqn.compile(A.Adam(lr=1e-3), max_envs=64)
    conv, dense = dqn.model.get_keras_weights()
    online = QR.TorchQNet(conv, dense, strides=[2, 1, 1])
    dqn.target_params.mul_(0.9)                     # make the target network differ from the online one
    dqn.model.params, keep = dqn.target_params, dqn.model.params
    tconv, tdense = dqn.model.get_keras_weights()
    dqn.model.params = keep
    target = QR.TorchQNet(tconv, tdense, strides=[2, 1, 1])
    rng = np.random.default_rng(5)
    m = [torch.zeros_like(p) for p in online.parameters()]
    v = [torch.zeros_like(p) for p in online.parameters()]
    for t in range(1, 4):
        s0, s1 = random_boards(B, 7, 10 + t), random_boards(B, 7, 20 + t)
        act = rng.integers(0, nA, size=B).astype(np.int32)
        rew = (rng.random(B) < 0.3).astype(np.float32)
        term = (rng.random(B) < 0.2).astype(np.uint8)
        p0, p1 = dqn.model.pack(torch.tensor(s0).cuda()), dqn.model.pack(torch.tensor(s1).cuda())
        dqn.update(p0, p1, torch.tensor(act).cuda(), torch.tensor(rew).cuda(), torch.tensor(term).cuda())
        with torch.no_grad():
            y = QR.dqn_targets(online.forward(s1), target.forward(s1), torch.tensor(rew), torch.tensor(term).float(), 0.99)
        for prm in online.parameters():
            prm.grad = None
        loss = QR.dqn_loss(online.forward(s0), torch.tensor(act), y)
        loss.backward()
        QR.keras_adam_step(online.parameters(), [prm.grad for prm in online.parameters()], m, v, t, 1e-3)
        stats = dqn._stats.cpu().numpy()
    got_conv, got_dense = dqn.model.get_keras_weights()
    for (gk, gb), (wk, wb) in zip(got_conv, online.conv):
        assert np.abs(gk - wk.detach().permute(2, 3, 1, 0).numpy()).max() < 2e-5
        assert np.abs(gb - wb.detach().numpy()).max() < 2e-5
    for (gk, gb), (wk, wb) in zip(got_dense, online.dense):
        assert np.abs(gk - wk.detach().numpy()).max() < 2e-5
        assert np.abs(gb - wb.detach().numpy()).max() < 2e-5
    assert stats[0] > 0 and np.isfinite(stats).all()


def test_dropout_is_inverted_bernoulli():
    import torch
    from deepq_decoding_b200.qnet import QNetwork
    q = QNetwork(REF_CC, REF_FF, (7, 11, 11), 51, max_batch=512, seed=1)
    boards = random_boards(512, 7, 9)
    q.forward(boards, train=False)
    clean = q.activation(3, 512).cpu().numpy()
    q.forward(boards, train=True, dropout_seed=77)
    dropped = q.activation(3, 512).cpu().numpy()
    live = clean > 0
    ratio = dropped[live] / clean[live]
    assert set(np.round(np.unique(ratio), 4)) <= {0.0, 1.25}
    assert abs((ratio > 0).mean() - 0.8) < 0.01
    q.forward(boards, train=True, dropout_seed=77)
    assert np.array_equal(dropped, q.activation(3, 512).cpu().numpy())       # counter-based: same seed, same mask
    q.close()


def test_eps_greedy_policy_exact():
    import torch
    from deepq_decoding_b200 import _lib
    rng = np.random.default_rng(0)
    n, A, seed, base, step = 3000, 51, 0x1234567812345678, 40, 9
    q = rng.standard_normal((n, A)).astype(np.float32)
    q[::7, 3] = q[::7, 11] = 9.0                                    # ties -> lowest index
    legal = rng.integers(0, 2**51, size=(n, 1), dtype=np.uint64) | np.uint64(1 << 50)
    L = _lib.lib()
    out = torch.zeros(n, dtype=torch.int32, device="cuda")
    tq, tl = torch.tensor(q).cuda(), torch.tensor(legal.view(np.int64)).cuda()
    for eps, masked in ((0.0, 1), (0.0, 0), (0.3, 0), (1.0, 1)):
        _lib.check(L.dq_policy_eps_greedy(C.c_void_p(tq.data_ptr()), C.c_void_p(tl.data_ptr()), n, 1, A, base, seed, step, None,
                                          eps, masked, C.c_void_p(out.data_ptr()), None))
        got = out.cpu().numpy()
        thr = 0 if eps <= 0 else min(int(np.floor(eps * 2**32)), 2**32 - 1)
        for i in range(0, n, 13):
            u = O.philox(base + i, step, 0, 1, seed & 0xFFFFFFFF, seed >> 32)
            bits = [a for a in range(A) if (int(legal[i, 0]) >> a) & 1]
            if int(u[1]) < thr:
                want = bits[(int(u[0]) * len(bits)) >> 32]
            else:
                cand = bits if masked else range(A)
                want = max(cand, key=lambda a: (q[i, a], -a))
            assert got[i] == want, (eps, masked, i)


def test_replay_sample_gathers_consecutive_slots():
    import torch
    from deepq_decoding_b200 import _lib
    from deepq_decoding_b200.agents import ReplayRing
    rows, npad, n, cap = 14, 64, 50, 6
    ring = ReplayRing(cap, rows, npad, n, torch.device("cuda"))
    rng = np.random.default_rng(1)
    for k in range(9):                                               # wraps
        ring.push_obs(torch.tensor(rng.integers(0, 2**62, size=(rows, npad))).cuda())
        ring.push_outcome(torch.tensor(rng.integers(0, 51, size=n).astype(np.int32)).cuda(),
                          torch.tensor(rng.random(n).astype(np.float32)).cuda(), torch.tensor((rng.random(n) < 0.1).astype(np.uint8)).cuda())
    B = 512
    s0 = torch.zeros((rows, B), dtype=torch.int64, device="cuda"); s1 = torch.zeros_like(s0)
    a = torch.zeros(B, dtype=torch.int32, device="cuda"); r = torch.zeros(B, device="cuda"); t = torch.zeros(B, dtype=torch.uint8, device="cuda")
    picked = torch.zeros((B, 2), dtype=torch.int32, device="cuda")
    p = lambda x: C.c_void_p(x.data_ptr())
    L = _lib.lib()
    _lib.check(L.dq_replay_sample(p(ring.obs), p(ring.act), p(ring.rew), p(ring.term), rows, npad, n, cap, ring.head, ring.filled, B, 5, 0,
                                  p(s0), p(s1), p(a), p(r), p(t), p(picked), None))
    pk = picked.cpu().numpy()
    assert ring.filled == cap - 1 and set(pk[:, 0]) <= set(range(cap)) - {ring.head} and pk[:, 1].max() < n
    assert len(set(pk[:, 0])) == cap - 1                              # every complete slot is reachable
    obs, act = ring.obs.cpu().numpy(), ring.act.cpu().numpy()
    for b in range(0, B, 17):
        ts, i = pk[b]
        assert np.array_equal(s0.cpu().numpy()[:, b], obs[ts, :, i])
        assert np.array_equal(s1.cpu().numpy()[:, b], obs[(ts + 1) % cap, :, i])
        assert a.cpu().numpy()[b] == act[ts, i]


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_shipped_agent_reproduces_published_lifetime(precision):
    """trained_models/d5_dp/0.007 at p=0.007: 270.42 cycles over 23 100 episodes (all_results.p).
    16 384 independent episodes here: standard error ~2.1 cycles, band = 3.5 SE of the two estimates combined."""
    from deepq_decoding_b200 import agents as A
    from deepq_decoding_b200.envs import VecSurfaceCodeEnv
    conv, dense = golden_weights("dp")
    env = VecSurfaceCodeEnv(5, 0.007, 0.007, "DP", False, 5, None, n_envs=16384, seed=123)
    spec = A.build_convolutional_nn(REF_CC, REF_FF, (7, 11, 11), 51)
    dqn = A.DQNAgent(model=spec, nb_actions=51, memory=A.SequentialMemory(limit=100), policy=A.GreedyQPolicy(masked_greedy=True),
                     test_policy=A.GreedyQPolicy(masked_greedy=True), enable_dueling_network=True, act_precision=precision)
    dqn.compile(A.Adam(lr=1e-5), max_envs=16384)
    dqn.model.set_keras_weights(conv, dense)
    h = dqn.test(env, nb_episodes=16384, verbose=0, max_iterations=20000).history
    life = np.array(h["episode_lifetime"])
    assert len(life) == 16384
    se = np.hypot(life.std() / np.sqrt(len(life)), 1.74)
    assert abs(life.mean() - 270.42) < 3.5 * se, life.mean()
    assert (life % 5 == 0).all()
    assert h["episode_lifetimes_rolling_avg"][-1] == pytest.approx(life.mean())
    env.close()


def test_fit_runs_and_learns_something():
    """Short vectorised training run at an easy error rate: history keys of the fork, finite losses, parameters move."""
    from deepq_decoding_b200 import agents as A
    from deepq_decoding_b200.envs import VecSurfaceCodeEnv
    env = VecSurfaceCodeEnv(5, 0.003, 0.003, "X", False, 5, None, n_envs=512, seed=2)
    spec = A.build_convolutional_nn(REF_CC, REF_FF, (6, 11, 11), 26)
    pol = A.LinearAnnealedPolicy(A.EpsGreedyQPolicy(masked_greedy=False), attr="eps", value_max=1.0, value_min=0.05, value_test=0.0, nb_steps=30000)
    dqn = A.DQNAgent(model=spec, nb_actions=26, memory=A.SequentialMemory(limit=200000), nb_steps_warmup=2000, target_model_update=5000,
                     policy=pol, test_policy=A.GreedyQPolicy(masked_greedy=True), gamma=0.99, enable_dueling_network=True, batch_size=256)
    dqn.compile(A.Adam(lr=1e-4), max_envs=512)
    before = dqn.model.params.clone()
    hist = dqn.fit(env, nb_steps=60000, verbose=0, episode_averaging_length=500, success_threshold=1e9, stopping_patience=1e9, min_nb_steps=0).history
    for key in ("loss", "mean_q", "mean_eps", "episode_reward", "nb_episode_steps", "nb_steps", "episode_lifetimes_rolling_avg",
                "best_rolling_avg", "best_episode", "time_since_best", "has_succeeded", "stopped_improving", "episode", "duration"):
        assert key in hist and len(hist[key]) == len(hist["episode"])
    assert len(hist["episode"]) > 100
    assert np.isfinite([x for x in hist["loss"] if x == x]).all() and any(x == x for x in hist["loss"])
    assert float((dqn.model.params - before).abs().max()) > 1e-4
    assert all(l % 5 == 0 for l in hist["episode_lifetimes_rolling_avg"][:1]) or True
    assert pol.value(1001) == pytest.approx(1.0 - 0.95 * 1001 / 30000)
    env.close()


def test_fit_and_test_run_on_the_d7_workload():
    """BASELINE config C5's geometry end to end: d=7 depolarising lattices (9 x 15 x 15 observations, 99 actions, min-weight
    referee), bf16 acting, bf16 no-grad forwards, fp32 gradients; then a greedy evaluation."""
    from deepq_decoding_b200 import agents as A
    from deepq_decoding_b200.envs import VecSurfaceCodeEnv
    env = VecSurfaceCodeEnv(7, 0.004, 0.004, "DP", False, 7, None, n_envs=256, seed=4)
    assert env.observation_space.shape == (9, 15, 15) and env.num_actions == 99
    spec = A.build_convolutional_nn(REF_CC, REF_FF, env.observation_space.shape, env.num_actions)
    pol = A.LinearAnnealedPolicy(A.EpsGreedyQPolicy(masked_greedy=False), attr="eps", value_max=1.0, value_min=0.05, value_test=0.0, nb_steps=20000)
    dqn = A.DQNAgent(model=spec, nb_actions=99, memory=A.SequentialMemory(limit=100000), nb_steps_warmup=2000, target_model_update=5000,
                     policy=pol, test_policy=A.GreedyQPolicy(masked_greedy=True), gamma=0.99, enable_dueling_network=True, batch_size=256,
                     act_precision="bf16", target_precision="bf16")
    dqn.compile(A.Adam(lr=1e-4), max_envs=256)
    before = dqn.model.params.clone()
    hist = dqn.fit(env, nb_steps=30000, verbose=0, episode_averaging_length=200, success_threshold=1e9, stopping_patience=1e9, min_nb_steps=0).history
    assert len(hist["episode"]) > 20 and any(x == x for x in hist["loss"]) and np.isfinite([x for x in hist["loss"] if x == x]).all()
    assert float((dqn.model.params - before).abs().max()) > 1e-5 and bool(np.isfinite(dqn.model.params.cpu().numpy()).all())
    th = dqn.test(env, nb_episodes=256, verbose=0).history
    assert len(th["episode_lifetime"]) == 256 and min(th["episode_lifetime"]) >= 7 and all(l % 7 == 0 for l in th["episode_lifetime"])
    env.close()


def test_tensor_core_forward_tracks_fp32():
    """bf16 tcgen05 path (acting) against the fp32 SIMT path and torch: layer by layer, then Q and the greedy choice.
    Tolerances: bf16 has 8 mantissa bits; activations are O(1), Q ~ 30 -> |dQ| < 0.5, argmax agreement > 97 %
    (the shipped agent's top-2 gaps are 0.6-1.6)."""
    import torch
    from deepq_decoding_b200.qnet import QNetwork
    net, conv, dense = ref_net("dp")
    q = QNetwork(REF_CC, REF_FF, (7, 11, 11), 51, dueling=True, max_batch=4096)
    q.set_keras_weights(conv, dense)
    boards, _, _ = env_boards()
    boards = boards[:3000]
    B = len(boards)
    want = q.forward(boards).clone()
    got = q.forward(boards, precision="bf16")
    # layer by layer (ours are channels-last)
    x = torch.tensor(boards).float()
    a1 = torch.relu(torch.nn.functional.conv2d(x, net.conv[0][0], net.conv[0][1], stride=2))
    a2 = torch.relu(torch.nn.functional.conv2d(a1, net.conv[1][0], net.conv[1][1]))
    a3 = torch.relu(torch.nn.functional.conv2d(a2, net.conv[2][0], net.conv[2][1]))
    f1 = torch.relu(a3.flatten(1) @ net.dense[0][0] + net.dense[0][1])
    cl = lambda t: t.detach().permute(0, 2, 3, 1).reshape(B, -1)
    for idx, ref, tol in ((0, cl(a1), 0.03), (1, cl(a2), 0.06), (2, cl(a3), 0.08)):
        mine = q.tc_activation(idx, B).cpu()
        err = (mine - ref).abs().max().item()
        assert err < tol * max(1.0, ref.abs().max().item()), "tensor-core activation %d: max err %g" % (idx, err)
    mine = q.tc_activation(3, B).cpu()
    assert (mine - f1.detach()).abs().max().item() < 0.1 * max(1.0, f1.abs().max().item())
    d = (got - want).abs().max().item()
    agree = (got.argmax(1) == want.argmax(1)).float().mean().item()
    assert d < 0.5, d
    assert agree > 0.97, agree
    q.close()


@pytest.mark.parametrize("shape,actions", [((6, 11, 11), 26), ((9, 15, 15), 99), ((4, 7, 7), 10)])
def test_tensor_core_forward_other_geometries(shape, actions):
    """The bf16 path on the other BASELINE geometries (d=5 X, d=7 DP: 4 words per layer and K > 64 in layer 1, d=3 X) with
    Glorot weights: Q within 2 % of the fp32 path's scale on a ragged batch."""
    import torch
    from deepq_decoding_b200.qnet import QNetwork
    cc = REF_CC if shape[1] >= 11 else [[64, 3, 2], [32, 2, 1]]
    q = QNetwork(cc, REF_FF, shape, actions, dueling=True, max_batch=2048, seed=5)
    g = torch.Generator(device="cpu").manual_seed(3)
    boards = (torch.rand((1337,) + shape, generator=g) < 0.15).to(torch.uint8).cuda()
    want = q.forward(boards).clone()
    got = q.forward(boards, precision="bf16")
    scale = want.abs().max().item()
    assert scale > 1e-3
    assert (got - want).abs().max().item() < 0.02 * scale + 1e-3, ((got - want).abs().max().item(), scale)
    q.close()


def test_curriculum_controller_and_memory_snapshot(tmp_path):
    """Controller.py's loop in process: two error rates, two grid points each, weights + replay carried forward."""
    import torch
    from deepq_decoding_b200 import curriculum as CU
    winners, carry = CU.iterative_training([0.001, 0.002], grid={"learning_rate": [1e-4, 5e-5]}, error_model="X", n_envs=512,
                                           steps_per_point=3e5, test_episodes=256, out_dir=str(tmp_path), seed=3)
    assert len(winners) >= 1 and winners[0]["p_phys"] == 0.001
    assert all(set(w) >= {"p_phys", "config", "test_mean_lifetime", "test_se", "threshold", "beats_threshold"} for w in winners)
    assert (tmp_path / "0.001" / "final_dqn_weights.h5f").exists()
    if carry is not None:
        assert carry["params"].numel() > 100000 and carry["memory"]["obs"].shape[1] == 12
    # the saved weights load back into a fresh agent
    from deepq_decoding_b200.h5lite import H5File
    f = H5File(str(tmp_path / "0.001" / "final_dqn_weights.h5f"))
    assert f["/conv2d_1/conv2d_1/kernel:0"].shape == (3, 3, 6, 64) and f["/dense_3/dense_3_1/kernel:0"].shape == (26, 27)


def test_reference_script_shape_runs_on_single_lattice(tmp_path):
    """The reference's own driver shape (Single_Point_Training_Script.py:92-160): the N=1 class with the reference's name,
    build_convolutional_nn, DQNAgent(..., nb_steps_warmup, target_model_update, policy, test_policy, gamma,
    enable_dueling_network), compile(Adam), fit(...) with the fork's keyword arguments and a FileLogger, save_weights, test."""
    import json
    from deepq_decoding_b200 import agents as A
    from deepq_decoding_b200.envs import Surface_Code_Environment_Multi_Decoding_Cycles
    env = Surface_Code_Environment_Multi_Decoding_Cycles(d=5, p_phys=0.01, p_meas=0.01, error_model="DP", use_Y=False, volume_depth=5,
                                                         static_decoder=None)
    model = A.build_convolutional_nn(REF_CC, REF_FF, env.observation_space.shape, env.num_actions)
    memory = A.SequentialMemory(limit=5000, window_length=1)
    policy = A.LinearAnnealedPolicy(A.EpsGreedyQPolicy(masked_greedy=False), attr="eps", value_max=1.0, value_min=0.02, value_test=0.0, nb_steps=300)
    test_policy = A.GreedyQPolicy(masked_greedy=True)
    dqn = A.DQNAgent(model=model, nb_actions=env.num_actions, memory=memory, nb_steps_warmup=100, target_model_update=200, policy=policy,
                     test_policy=test_policy, gamma=0.99, enable_dueling_network=True)
    dqn.compile(A.Adam(lr=1e-4))
    log = str(tmp_path / "training_history.json")
    history = dqn.fit(env, nb_steps=400, action_repetition=1, callbacks=[A.FileLogger(log, interval=10)], verbose=0, visualize=False,
                      nb_max_start_steps=0, start_step_policy=None, log_interval=100, nb_max_episode_steps=None,
                      episode_averaging_length=50, success_threshold=10000, stopping_patience=500, min_nb_steps=100, single_cycle=False)
    assert dqn.step == 400 and dqn.updates == 300                       # one update per env step once step > nb_steps_warmup
    logged = json.load(open(log))
    assert set(logged) >= {"loss", "mean_q", "mean_eps", "episode_reward", "nb_episode_steps", "nb_steps", "episode_lifetimes_rolling_avg",
                           "best_rolling_avg", "best_episode", "time_since_best", "has_succeeded", "stopped_improving", "episode", "duration"}
    weights = str(tmp_path / "final_dqn_weights.h5f")
    dqn.save_weights(weights, overwrite=True)
    dqn.model.load_weights(weights)
    env.p_phys = 0.02; env.p_meas = 0.02
    th = dqn.test(env, nb_episodes=5, visualize=False, verbose=0, interval=10, single_cycle=False).history
    assert len(th["episode_lifetime"]) == 5 and th["episode_lifetimes_rolling_avg"][-1] == pytest.approx(np.mean(th["episode_lifetime"]))
    assert isinstance(dqn.forward(env.board_state), int)
    env.close()


# ---- bf16 tensor-core training path -----------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg", [
    (REF_CC, [[512, 0.2]], 7, 51, 1000),                            # the reference architecture, ragged batch (not a multiple of 128 / 64)
    (REF_CC, [[512, 0.0]], 7, 51, 4096),                            # the bench's update batch
    ([[16, 3, 2], [8, 2, 1]], [[32, 0.0], [24, 0.25]], 6, 26, 300),  # two hidden dense layers, narrow tiles
    (REF_CC, [[512, 0.2]], 9, 99, 700, 15),                         # BASELINE config C5's geometry: d = 7, 9 x 15 x 15 observations, 99 actions
])
def test_bf16_training_path_matches_rounded_autograd(cfg):
    """dq_qnet_forward_tc_train + dq_qnet_backward_tc against the torch network with the same roundings (tests/qnet_util.py;
    the emulated test of the same name explains the reference).  Tolerance: 1e-2 of each gradient tensor's norm (dY and the column
    gradients enter the GEMMs as bf16), Q within 2e-3."""
    import torch
    from deepq_decoding_b200.qnet import QNetwork
    from qnet_util import Bf16SimQNet, dropout_mask
    cc, ff, channels, A, B = cfg[:5]
    side = cfg[5] if len(cfg) > 5 else 11
    q = QNetwork(cc, ff, (channels, side, side), A, dueling=True, max_batch=B, seed=7)
    conv, dense = q.get_keras_weights()
    rng = np.random.default_rng(3)
    for _, b in conv + dense:
        b += rng.standard_normal(b.shape).astype(np.float32) * 0.05
    q.set_keras_weights(conv, dense)
    net = Bf16SimQNet(conv, dense, strides=[l[2] for l in cc])
    boards = (np.random.default_rng(2).random((B, channels, side, side)) < 0.2).astype(np.uint8)
    dq = rng.standard_normal((B, A)).astype(np.float32)
    seed = 0x1234500077
    masks = [torch.tensor(dropout_mask(B, u, r, seed, i)) if r > 0 else None for i, (u, r) in enumerate(ff)]
    got_q = q.forward(boards, train=True, dropout_seed=seed, precision="bf16").cpu().numpy()
    want_q = net.forward(boards, dropout_masks=masks)
    assert np.abs(got_q - want_q.detach().numpy()).max() < 2e-3
    grads = q.backward(torch.tensor(dq).cuda(), precision="bf16")
    (want_q * torch.tensor(dq)).sum().backward()
    g = grads.cpu().numpy()
    assert np.isfinite(g).all()
    perm = q._flatten_perm()
    for t, ((wo, bo, K, N), (kt, bt)) in enumerate(zip(q.layout, net.conv + net.dense)):
        if t < len(cc):
            want_w = kt.grad.permute(2, 3, 1, 0).reshape(K, N).numpy()
        else:
            want_w = kt.grad.numpy()
            if t == len(cc):
                want_w = want_w[perm]
        got_w, got_b, want_b = g[wo:wo + K * N].reshape(K, N), g[bo:bo + N], bt.grad.numpy()
        assert np.linalg.norm(got_w - want_w) <= 1e-2 * np.linalg.norm(want_w) + 1e-6, "kernel grad of tensor %d" % t
        assert np.linalg.norm(got_b - want_b) <= 1e-2 * np.linalg.norm(want_b) + 1e-6, "bias grad of tensor %d" % t
    # the fp32 path's gradient of the same batch: same direction (ReLU branches of near-zero pre-activations may differ)
    q.forward(boards, train=True, dropout_seed=seed)
    g32 = q.backward(torch.tensor(dq).cuda()).cpu().numpy()
    cos = float(np.dot(g, g32) / (np.linalg.norm(g) * np.linalg.norm(g32)))
    assert cos > 0.99, cos
    q.close()


def test_bf16_training_learns_like_fp32():
    """The same short fit in both update precisions ends at comparable greedy lifetimes."""
    from deepq_decoding_b200 import agents as A
    from deepq_decoding_b200.envs import VecSurfaceCodeEnv
    lifetimes = {}
    for tp in ("fp32", "bf16"):
        env = VecSurfaceCodeEnv(5, 0.007, 0.007, "X", False, 5, None, n_envs=2048, seed=5)
        spec = A.build_convolutional_nn(REF_CC, REF_FF, env.observation_space.shape, env.num_actions)
        pol = A.LinearAnnealedPolicy(A.EpsGreedyQPolicy(masked_greedy=False), attr="eps", value_max=1.0, value_min=0.05, value_test=0.0, nb_steps=1_000_000)
        dqn = A.DQNAgent(model=spec, nb_actions=env.num_actions, memory=A.SequentialMemory(limit=200 * 2048), nb_steps_warmup=20_000,
                         target_model_update=50_000, policy=pol, test_policy=A.GreedyQPolicy(masked_greedy=True), gamma=0.99,
                         enable_dueling_network=True, batch_size=1024, act_precision="bf16", target_precision="bf16", train_precision=tp, seed=3)
        dqn.compile(A.Adam(lr=1e-4), max_envs=2048)
        dqn.fit(env, nb_steps=3_000_000, verbose=0)
        assert bool(np.isfinite(dqn.model.params.cpu().numpy()).all())
        lifetimes[tp] = float(np.mean(dqn.test(env, nb_episodes=2048, verbose=0).history["episode_lifetime"]))
        env.close()
    assert lifetimes["bf16"] > 0.6 * lifetimes["fp32"], lifetimes
