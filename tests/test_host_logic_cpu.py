"""CPU tests of the host-side logic and of the Q-network restatement in oracle/ (no GPU, no compute through the C ABI):
the torch restatement against the reference's published notebook output, the HDF5 subset reader/writer, referee tables,
exploration schedule."""
import os

import numpy as np
import pytest

from qnet_util import golden_weights

HERE = os.path.dirname(os.path.abspath(__file__))


def test_torch_restatement_reproduces_the_notebook_decode():
    """README.md:719-829 (notebook 3): the shipped d5_x/0.007 agent on that volume picks correction 21, then the identity with
    Q = 34.91.  Pins oracle/qnet_ref.py (layouts, Flatten order, dueling head) without a GPU."""
    import torch
    from oracle import qnet_ref as R
    conv, dense = golden_weights("x")
    net = R.TorchQNet(conv, dense, [2, 1, 1], dueling=True)
    state = np.load(os.path.join(HERE, "golden", "qnet_d5_x_kat.npz"))["state0"].astype(np.float32)
    state[5] = 0
    with torch.no_grad():
        q0 = net.forward(state[None])[0].numpy()
    legal_first = int(np.argmax(q0[:25]))                # greedy over qubit actions: the notebook's first correction
    assert legal_first == 21 and abs(float(q0[21]) - 35.16) < 0.02
    state[5, 2 * (21 // 5) + 1, 2 * (21 % 5) + 1] = 1
    with torch.no_grad():
        q1 = net.forward(state[None])[0].numpy()
    assert int(np.argmax(q1)) == 25 and abs(float(q1[25]) - 34.91) < 0.01


def test_keras_adam_and_targets_restatement():
    import torch
    from oracle import qnet_ref as R
    p, g = [torch.tensor([1.0, -2.0])], [torch.tensor([0.5, 0.25])]
    m, v = [torch.zeros(2)], [torch.zeros(2)]
    R.keras_adam_step(p, g, m, v, 1, lr=0.1)
    # t = 1: m = 0.1 g, v = 0.001 g^2, lr_t = lr sqrt(1-b2)/(1-b1)  ->  step = lr g / (|g| + eps sqrt(...)) ~ lr sign(g)
    assert torch.allclose(p[0], torch.tensor([0.9, -2.1]), atol=1e-5)
    y = R.dqn_targets(torch.tensor([[1.0, 3.0], [2.0, 0.0]]), torch.tensor([[10.0, 20.0], [30.0, 40.0]]), torch.tensor([1.0, 0.0]),
                      torch.tensor([0.0, 1.0]), 0.5)
    assert torch.allclose(y, torch.tensor([1.0 + 0.5 * 20.0, 0.0]))          # argmax by the online net, value by the target net
    loss = R.dqn_loss(torch.tensor([[1.0, 2.0], [3.0, 4.0]]), torch.tensor([1, 0]), torch.tensor([4.0, 3.0]))
    assert float(loss) == pytest.approx(0.5 * ((4.0 - 2.0) ** 2 + 0.0) / 2)


def test_h5_subset_round_trip(tmp_path):
    from deepq_decoding_b200 import h5lite
    rng = np.random.default_rng(0)
    tree = {"@backend": "tensorflow", "@layer_names": np.array([b"conv2d_1", b"dense_1"]),
            "conv2d_1": {"@weight_names": np.array([b"conv2d_1/kernel:0", b"conv2d_1/bias:0"]),
                         "conv2d_1": {"kernel:0": rng.normal(size=(3, 3, 7, 4)).astype(np.float32), "bias:0": np.zeros(4, np.float32)}},
            "dense_1": {"dense_1_1": {"kernel:0": rng.normal(size=(5, 6)).astype(np.float32), "bias:0": np.arange(6, dtype=np.float32)}}}
    path = str(tmp_path / "w.h5f")
    h5lite.write_h5(path, tree)
    f = h5lite.H5File(path)
    assert sorted(f.keys("/")) == ["conv2d_1", "dense_1"]
    assert np.array_equal(f["/conv2d_1/conv2d_1/kernel:0"], tree["conv2d_1"]["conv2d_1"]["kernel:0"])
    assert np.array_equal(f["/dense_1/dense_1_1/bias:0"], np.arange(6, dtype=np.float32))
    assert f.attrs("/")["backend"] in ("tensorflow", b"tensorflow")
    names = [n.decode() if isinstance(n, bytes) else n for n in f.attrs("/conv2d_1")["weight_names"]]
    assert names == ["conv2d_1/kernel:0", "conv2d_1/bias:0"]
    with pytest.raises(ValueError):
        bad = tmp_path / "bad.h5"
        bad.write_bytes(b"not hdf5 at all")
        h5lite.H5File(str(bad))
    # structural checks libhdf5 applies when it opens a group (no h5py in this image to open the file with): every local heap
    # (H5HL prefix) must carry a free-list head that is H5HL_FREE_NULL (= 1) or an offset inside its data segment, and a data
    # segment that lies inside the file
    heaps = _local_heaps(open(path, "rb").read())
    assert len(heaps) == 5                     # root, conv2d_1, conv2d_1/conv2d_1, dense_1, dense_1/dense_1_1
    for size, free_head, seg_addr, file_len in heaps:
        assert free_head == 1 or free_head < size, "libhdf5: 'bad heap free list'"
        assert size >= 16 and size % 8 == 0 and seg_addr + size <= file_len


def _local_heaps(buf):
    """(data segment size, free-list head, data segment address, file length) of every local heap in an HDF5 image."""
    import struct
    out, pos = [], buf.find(b"HEAP")
    while pos >= 0:
        if pos % 8 == 0 and buf[pos + 4] == 0:
            size, free_head, seg = struct.unpack_from("<QQQ", buf, pos + 8)
            out.append((size, free_head, seg, len(buf)))
        pos = buf.find(b"HEAP", pos + 4)
    return out


@pytest.mark.reference
def test_reference_written_heaps_satisfy_the_same_rule():
    """The check above, applied to files h5py wrote (the reference's shipped weights): validates the check itself."""
    import glob
    files = glob.glob("/root/reference/trained_models/d5_dp/0.007/final_dqn_weights.h5f") + glob.glob("/root/reference/example_notebooks/referee_decoders/nn_d5_X_p5")
    assert files
    for f in files:
        heaps = _local_heaps(open(f, "rb").read())
        assert heaps
        for size, free_head, seg_addr, file_len in heaps:
            assert free_head == 1 or free_head < size
            assert seg_addr + size <= file_len


@pytest.mark.parametrize("d,model", [(3, "X"), (3, "DP"), (5, "X")])
def test_minimum_weight_referee_corrects_single_errors(d, model):
    """The referee tables built here for the distances the reference ships none for: the zero syndrome maps
    to the trivial class, and for every single-qubit error the table names the homology class of that error itself (the
    minimum-weight explanation of its syndrome), so the env's `referee class != true class` test does not end the episode."""
    from deepq_decoding_b200 import referee as REF
    from deepq_decoding_b200.envs import true_syndrome_of
    from oracle import oracle as O
    orc = O.OracleVecEnv(d, model, False, d, 0.01, 0.01, 1, 0)
    lut = REF.min_weight(d, model)
    zero = np.zeros((d + 1, d + 1), np.int8)
    assert lut.classify(zero) == 0
    assert np.array_equal(lut.predict(zero.reshape(1, -1)), np.eye(lut.n_classes, dtype=np.float32)[:1])
    paulis = (1,) if model == "X" else (1, 2, 3)
    for r in range(d):
        for c in range(d):
            for pl in paulis:
                hidden = np.zeros((d, d), np.int8)
                hidden[r, c] = pl
                syn, label = orc.syndrome_of(hidden)
                assert np.array_equal(syn, true_syndrome_of(hidden))
                assert lut.classify(syn) == label, (r, c, pl)
    packed = REF.pack2(np.array([0, 1, 2, 3, 3, 2, 1], np.uint8))
    assert list(REF.unpack2(packed, 7)) == [0, 1, 2, 3, 3, 2, 1]
    saved = os.path.join(os.path.dirname(REF.__file__), "data", "referee_d%d_%s.lut" % (d, model))
    shipped = REF.RefereeLUT.load(saved)
    if d != 5:                                           # d = 5 ships the reference's own Keras referees, tabulated
        assert np.array_equal(shipped.lut_a, lut.lut_a)  # the committed table is this construction
    else:
        agree = np.mean(REF.unpack2(shipped.lut_a, 1 << 12) == REF.unpack2(lut.lut_a, 1 << 12))
        assert agree > 0.9                               # the trained referee and the minimum-weight one mostly agree


def test_exploration_schedule_and_config_objects():
    from deepq_decoding_b200 import agents as A
    pol = A.LinearAnnealedPolicy(A.EpsGreedyQPolicy(masked_greedy=False), attr="eps", value_max=1.0, value_min=0.02, value_test=0.0, nb_steps=100000)
    assert pol.value(0) == 1.0 and pol.value(50000) == pytest.approx(0.51) and pol.value(10 ** 7) == 0.02
    assert pol.value(123, training=False) == 0.0 and pol.masked_greedy is False
    assert A.GreedyQPolicy(masked_greedy=True).masked_greedy is True
    with pytest.raises(NotImplementedError):
        A.BoltzmannQPolicy()
    mem = A.SequentialMemory(limit=50000, window_length=1)
    assert mem.limit == 50000
    with pytest.raises(ValueError):
        A.SequentialMemory(limit=10, window_length=4)
    spec = A.build_convolutional_nn([[64, 3, 2]], [[512, 0.2]], (7, 11, 11), 51)
    assert spec.input_shape == (7, 11, 11) and spec.num_actions == 51


@pytest.mark.parametrize("L,thr,pat,min_steps", [(1000, 1e5, 1e9, 0), (7, 300, 5, 100), (1, 50, 2, 0), (50, 1e9, 30, 2000)])
def test_episode_book_batched_form_equals_per_episode_form(L, thr, pat, min_steps):
    """EpisodeBook.finish_many (one call per drain of the vectorised fit) against finish_episode one episode at a time, on random
    lifetimes in ragged batches (empty ones included), short windows and active stop rules: same entries, same state, to the bit."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("dq_episodes", os.path.join(os.path.dirname(HERE), "deepq_decoding_b200", "episodes.py"))
    ep = importlib.util.module_from_spec(spec); spec.loader.exec_module(ep)
    rng = np.random.default_rng(L)
    a, b, steps = ep.EpisodeBook(L, thr, pat, min_steps), ep.EpisodeBook(L, thr, pat, min_steps), 0
    for chunk in range(30):
        m = int(rng.integers(0, 300))
        life = (rng.geometric(0.02, size=m) * 5).astype(np.int64)
        nb = steps + np.cumsum(rng.integers(1, 50, size=m))
        steps = int(nb[-1]) if m else steps
        want = [a.finish_episode(int(x), int(y)) for x, y in zip(life, nb)]
        got = b.finish_many(life, nb)
        for k in got:
            assert np.asarray(got[k]).tolist() == [w[k] for w in want], (chunk, k)
        assert (a.stop, a.episode, a.best_avg, a.best_episode, a.win_n, a.win_sum) == (b.stop, b.episode, b.best_avg, b.best_episode, b.win_n, b.win_sum)
        assert np.array_equal(a.win, b.win)
