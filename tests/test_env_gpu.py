"""GPU parity of the CUDA environment (through the C ABI) against the CPU oracle.

Bit-exact on every integer output (obs, done, lifetime, legal mask, packed state) and on the
0/1 float reward, for every step of seeded trajectories, at sizes the oracle finishes in seconds.
Full-size (BASELINE config) runs use size-independent properties.
"""
import numpy as np
import pytest

from oracle import oracle as O
from deepq_decoding_b200 import referee as REF

pytestmark = pytest.mark.gpu


def random_referee(rng, d, model):
    ns = d * d - 1
    if model == "X" or d == 7:
        la = rng.integers(0, 256, size=(1 << (ns // 2)) // 4 + 1, dtype=np.uint8) & 0x55
        lb = rng.integers(0, 256, size=(1 << (ns // 2)) // 4 + 1, dtype=np.uint8) & 0x55
        return REF.RefereeLUT(d, model, REF.SPLIT, la, lb if model == "DP" else None)
    return REF.RefereeLUT(d, model, REF.JOINT, rng.integers(0, 256, size=(1 << ns) // 4 + 1, dtype=np.uint8))


def make_pair(d, model, use_Y, vd, p, n, seed, base=0, referee=None, auto_reset=True):
    from deepq_decoding_b200.envs import VecSurfaceCodeEnv
    rng = np.random.default_rng(1000 * d + vd)
    ref = referee or random_referee(rng, d, model)
    env = VecSurfaceCodeEnv(d, p, p, model, use_Y, vd, ref, n_envs=n, seed=seed, env_id_base=base, auto_reset=auto_reset)
    o = O.OracleVecEnv(d, model, use_Y, vd, p, p, n, seed, base)
    o.set_referee(ref.mode, ref.lut_a, ref.lut_b)
    return env, o


def compare_state(env, o, idx):
    for i in idx:
        st, ost = env.decode_state(i), o.get_env(i)
        assert np.array_equal(st["hidden_state"], ost["hidden"])
        assert st["lifetime"] == ost["lifetime"] and st["attempts"] == ost["attempts"] and st["done"] == ost["done"]


CASES = [(3, "X", False, 3, 0.05, 1),        # BASELINE config 1: one lattice
         (3, "X", False, 3, 0.05, 1000),
         (5, "X", False, 5, 0.02, 777),       # ragged tail CTA
         (5, "DP", False, 5, 0.02, 1024),
         (5, "DP", True, 3, 0.03, 33),
         (7, "DP", False, 7, 0.011, 300),
         (7, "DP", True, 8, 0.02, 40),        # deepest volume, widest action mask (148 actions, 3 words)
         (3, "DP", True, 2, 0.08, 64)]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "d%d_%s_y%d_vd%d_n%d" % (c[0], c[1], c[2], c[3], c[5]))
def test_trajectory_parity(case):
    import torch
    d, model, use_Y, vd, p, n = case
    env, o = make_pair(d, model, use_Y, vd, p, n, seed=42, base=5)
    rng = np.random.default_rng(d + n)
    obs = env.reset().cpu().numpy()
    oobs, olegal = o.reset()
    assert np.array_equal(obs, oobs)
    assert np.array_equal(env.legal_mask.cpu().numpy().view(np.uint64), olegal)
    steps = 150 if n <= 1024 else 40
    ndone = 0
    for t in range(steps):
        acts = env.random_legal_actions(t).cpu().numpy().copy()
        assert np.array_equal(acts, o.random_legal_actions(olegal, t)), "random-legal policy"
        arb = rng.random(n) < 0.2
        acts[arb] = rng.integers(-1, o.A + 1, size=int(arb.sum()))      # includes out-of-range -> identity
        oacts = np.where((acts < 0) | (acts >= o.A), o.A - 1, acts).astype(np.int32)
        obs, rew, done, info = env.step(torch.from_numpy(acts).cuda())
        oobs, orew, odone, olife, olegal = o.step(oacts, auto_reset=True)
        assert np.array_equal(obs.cpu().numpy(), oobs), "obs t=%d" % t
        assert np.array_equal(rew.cpu().numpy(), orew), "reward t=%d" % t
        assert np.array_equal(done.cpu().numpy(), odone.astype(bool)), "done t=%d" % t
        assert np.array_equal(info["lifetime"].cpu().numpy(), olife), "lifetime t=%d" % t
        assert np.array_equal(info["legal_mask"].cpu().numpy().view(np.uint64), olegal), "legal t=%d" % t
        ndone += int(odone.sum())
        if t % 50 == 0:
            compare_state(env, o, rng.integers(0, n, size=min(n, 4)))
    if n >= 64:
        assert ndone > 0
    env.close()


def test_no_auto_reset_keeps_reference_semantics():
    """auto_reset=0: done is sticky and the lattice keeps evolving, as the reference object does."""
    import torch
    env, o = make_pair(5, "DP", False, 5, 0.05, 256, seed=7, auto_reset=False)
    env.reset(); _, olegal = o.reset()
    for t in range(60):
        acts = o.random_legal_actions(olegal, t)
        obs, rew, done, info = env.step(torch.from_numpy(acts).cuda())
        oobs, orew, odone, olife, olegal = o.step(acts, auto_reset=False)
        assert np.array_equal(obs.cpu().numpy(), oobs)
        assert np.array_equal(done.cpu().numpy(), odone.astype(bool))
        assert np.array_equal(info["lifetime"].cpu().numpy(), olife)
    assert odone.sum() > 10
    # reset clears done / lifetime but keeps the position in the random stream
    obs = env.reset().cpu().numpy(); oobs, _ = o.reset()
    assert np.array_equal(obs, oobs)
    compare_state(env, o, [0, 100, 255])
    env.close()


def test_host_buffer_entry_points_and_noise_setter():
    env, o = make_pair(5, "X", False, 5, 0.01, 100, seed=3)
    obs, legal = env.reset_host(); oobs, olegal = o.reset()
    assert np.array_equal(obs, oobs) and np.array_equal(legal, olegal)
    for t in range(40):
        if t == 20:                       # the test sweep mutates env.p_phys / env.p_meas in place (SPTS:200-201)
            env.p_phys = 0.03; env.p_meas = 0.02; o.set_noise(0.03, 0.02)
        acts = o.random_legal_actions(olegal, t)
        obs, rew, done, info = env.step_host(acts)
        oobs, orew, odone, olife, olegal = o.step(acts, auto_reset=True)
        assert np.array_equal(obs, oobs) and np.array_equal(rew, orew) and np.array_equal(done, odone.astype(bool))
        assert np.array_equal(info["lifetime"], olife) and np.array_equal(info["legal_mask"], olegal)
    env.close()


def test_host_buffer_calls_with_pageable_caller_buffers():
    """The C ABI's *_host calls with ordinary (pageable) numpy buffers -- the library stages through its own pinned block -- against the
    binding's pinned buffers, which the kernel reads and writes directly: same outputs, both equal to the oracle."""
    import ctypes as C
    from deepq_decoding_b200 import _lib
    env, o = make_pair(5, "DP", False, 5, 0.02, 300, seed=9)
    raw, _ = make_pair(5, "DP", False, 5, 0.02, 300, seed=9)
    L = _lib.lib()
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    n, Wm = 300, env.mask_words
    obs = np.zeros((n,) + tuple(env.obs.shape[1:]), np.uint8); legal = np.zeros((n, Wm), np.uint64)
    rew = np.zeros(n, np.float32); done = np.zeros(n, np.uint8); life = np.zeros(n, np.int32)
    _lib.check(L.dq_env_reset_host(raw._h, vp(obs), vp(legal)))
    pobs, plegal = env.reset_host(); oobs, olegal = o.reset()
    assert np.array_equal(obs, oobs) and np.array_equal(legal, olegal) and np.array_equal(pobs, oobs)
    for t in range(25):
        acts = np.ascontiguousarray(o.random_legal_actions(olegal, t), dtype=np.int32)
        _lib.check(L.dq_env_step_host(raw._h, vp(acts), vp(obs), vp(rew), vp(done), vp(life), vp(legal), 1))
        pobs, prew, pdone, pinfo = env.step_host(acts)
        oobs, orew, odone, olife, olegal = o.step(acts, auto_reset=True)
        assert np.array_equal(obs, oobs) and np.array_equal(rew, orew) and np.array_equal(done, odone) and np.array_equal(life, olife)
        assert np.array_equal(legal, olegal)
        assert np.array_equal(pobs, oobs) and np.array_equal(prew, orew) and np.array_equal(pdone, odone.astype(bool))
        assert np.array_equal(pinfo["lifetime"], olife) and np.array_equal(pinfo["legal_mask"], olegal)
    env.close(); raw.close()


def test_sharding_is_invisible():
    """Lattices [0,N) on one handle == two handles covering [0,N/2) and [N/2,N) (stream id = global index)."""
    import torch
    full, _ = make_pair(5, "DP", False, 5, 0.02, 128, seed=11)
    lo, _ = make_pair(5, "DP", False, 5, 0.02, 64, seed=11, base=0)
    hi, _ = make_pair(5, "DP", False, 5, 0.02, 64, seed=11, base=64)
    a = full.reset().clone(); b = torch.cat([lo.reset(), hi.reset()])
    assert torch.equal(a, b)
    for t in range(30):
        fa = full.random_legal_actions(t).clone()
        la, ha = lo.random_legal_actions(t).clone(), hi.random_legal_actions(t).clone()
        assert torch.equal(fa, torch.cat([la, ha]))
        a = full.step(fa)[0]
        b = torch.cat([lo.step(la)[0], hi.step(ha)[0]])
        assert torch.equal(a, b)
    for e in (full, lo, hi):
        e.close()


def test_graph_capturable_policy_counter():
    """dq_policy_random_legal_next == dq_policy_random_legal at the device-side step index, which it advances."""
    import ctypes as C
    import torch
    from deepq_decoding_b200 import _lib
    env, _ = make_pair(5, "DP", False, 5, 0.02, 1000, seed=5)
    env.reset()
    L = _lib.lib()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    out = torch.zeros(1000, dtype=torch.int32, device="cuda")
    _lib.check(L.dq_policy_seek(env._h, 7, st))
    for step in (7, 8, 9):
        _lib.check(L.dq_policy_random_legal_next(env._h, C.c_void_p(env.legal_mask.data_ptr()), C.c_void_p(out.data_ptr()), st))
        assert torch.equal(out, env.random_legal_actions(step))
    env.close()


def test_fused_random_policy_step_equals_policy_then_step():
    """dq_env_step_random == dq_policy_random_legal_next followed by dq_env_step, bit for bit."""
    import ctypes as C
    import torch
    from deepq_decoding_b200 import _lib
    a, _ = make_pair(5, "DP", False, 5, 0.02, 1000, seed=21)
    b, _ = make_pair(5, "DP", False, 5, 0.02, 1000, seed=21)
    L = _lib.lib()
    p = lambda x: C.c_void_p(x.data_ptr())
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert torch.equal(a.reset(), b.reset())
    for env in (a, b):
        _lib.check(L.dq_policy_seek(env._h, 3, st))
    picks = torch.zeros(1000, dtype=torch.int32, device="cuda")
    acts = torch.zeros(1000, dtype=torch.int32, device="cuda")
    for t in range(60):
        _lib.check(L.dq_env_step_random(a._h, p(a.obs), p(a.reward), p(a.done), p(a.lifetime), p(a.legal_mask), p(picks), 1, st))
        _lib.check(L.dq_policy_random_legal_next(b._h, p(b.legal_mask), p(acts), st))
        _lib.check(L.dq_env_step(b._h, p(acts), p(b.obs), p(b.reward), p(b.done), p(b.lifetime), p(b.legal_mask), 1, st))
        assert torch.equal(picks, acts), t
        assert torch.equal(a.obs, b.obs) and torch.equal(a.legal_mask, b.legal_mask) and torch.equal(a.lifetime, b.lifetime), t
        assert torch.equal(a.done, b.done) and torch.equal(a.reward, b.reward)
    a.close(); b.close()


@pytest.mark.parametrize("d,model,n", [(5, "DP", 1000), (3, "X", 37), (7, "DP", 200), (5, "DP", 16384), (7, "DP", 8192)])
def test_rollout_equals_single_steps_and_oracle(d, model, n):
    """dq_env_rollout_random(n_steps) == n_steps x dq_env_step_random == the oracle stepping the same policy stream.
    The last two cases are the full per-GPU lattice counts of BASELINE configs C3 and C5, bit for bit against the oracle."""
    import ctypes as C
    import torch
    from deepq_decoding_b200 import _lib
    a, orc = make_pair(d, model, False, d, 0.03, n, seed=33)
    b, _ = make_pair(d, model, False, d, 0.03, n, seed=33)
    L = _lib.lib()
    p = lambda x: C.c_void_p(x.data_ptr())
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    obs0 = a.reset()
    assert torch.equal(obs0, b.reset())
    o_obs, o_legal = orc.reset()
    assert np.array_equal(obs0.cpu().numpy(), o_obs)
    S, slots = 23, 5
    ring = torch.zeros((slots,) + tuple(a.obs.shape), dtype=torch.uint8, device="cuda")
    step = 0
    for rnd in range(2):                                     # the second rollout continues the first (policy counter, ring cursor)
        first = (rnd * S) % slots
        out = a.rollout_random(S, ring, first_slot=first, keep=("reward", "done", "lifetime", "actions", "legal"))
        picks = torch.zeros(n, dtype=torch.int32, device="cuda")
        for s in range(S):
            _lib.check(L.dq_env_step_random(b._h, p(b.obs), p(b.reward), p(b.done), p(b.lifetime), p(b.legal_mask), p(picks), 1, st))
            assert torch.equal(out["actions"][s], picks), (rnd, s)
            assert torch.equal(out["reward"][s], b.reward) and torch.equal(out["done"][s], b.done) and torch.equal(out["lifetime"][s], b.lifetime)
            assert torch.equal(out["legal"][s], b.legal_mask), (rnd, s)
            if s >= S - slots:                               # the last `slots` steps are still in the ring
                assert torch.equal(ring[(first + s) % slots], b.obs), (rnd, s)
            # oracle on the same stream
            o_act = orc.random_legal_actions(o_legal, step)
            o_obs, o_rew, o_done, o_life, o_legal = orc.step(o_act)
            assert np.array_equal(picks.cpu().numpy(), o_act) and np.array_equal(b.obs.cpu().numpy(), o_obs)
            assert np.array_equal(out["lifetime"][s].cpu().numpy(), o_life) and np.array_equal(out["done"][s].cpu().numpy(), o_done)
            step += 1
        assert torch.equal(a.get_state_words(), b.get_state_words())
    a.close(); b.close()


def test_rollout_without_auto_reset_matches_single_steps():
    """auto_reset = 0: finished lattices stay finished inside a rollout exactly as across single-step calls."""
    import ctypes as C
    import torch
    from deepq_decoding_b200 import _lib
    a, _ = make_pair(5, "X", False, 5, 0.05, 300, seed=44, auto_reset=False)
    b, _ = make_pair(5, "X", False, 5, 0.05, 300, seed=44, auto_reset=False)
    L = _lib.lib()
    p = lambda x: C.c_void_p(x.data_ptr())
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert torch.equal(a.reset(), b.reset())
    S = 40
    out = a.rollout_random(S, None, keep=("reward", "done", "lifetime", "actions"))
    picks = torch.zeros(300, dtype=torch.int32, device="cuda")
    for s in range(S):
        _lib.check(L.dq_env_step_random(b._h, None, p(b.reward), p(b.done), p(b.lifetime), p(b.legal_mask), p(picks), 0, st))
        assert torch.equal(out["actions"][s], picks) and torch.equal(out["done"][s], b.done) and torch.equal(out["lifetime"][s], b.lifetime), s
    assert int(out["done"][-1].sum()) > 0                     # some lattices did finish, and stayed so
    assert torch.equal(a.get_state_words(), b.get_state_words())
    a.close(); b.close()


def test_state_roundtrip_and_injection():
    """get_state/set_state: a restored handle continues bit-identically (checkpoint contract)."""
    import torch
    env, _ = make_pair(5, "DP", False, 5, 0.02, 96, seed=5)
    env.reset()
    for t in range(10):
        env.step(env.random_legal_actions(t))
    words = env.get_state_words().clone()
    legal = env.legal_mask.clone()
    ref_traj = []
    for t in range(10, 20):
        ref_traj.append(env.step(env.random_legal_actions(t))[0].clone())
    env.set_state_words(words)
    env.legal_mask.copy_(legal)
    for k, t in enumerate(range(10, 20)):
        assert torch.equal(env.step(env.random_legal_actions(t))[0], ref_traj[k])
    env.close()


def test_single_error_syndrome_kat():
    """Notebook 3 cells 18/20 (README.md:719-780): X on qubit (4,1) of a d=5 lattice lights true
    syndrome bits (4,1) and (5,2)."""
    import torch
    from deepq_decoding_b200.envs import VecSurfaceCodeEnv, true_syndrome_of
    hidden = np.zeros((5, 5), int); hidden[4, 1] = 1
    syn = true_syndrome_of(hidden)
    assert {(int(a), int(b)) for a, b in zip(*np.nonzero(syn))} == {(4, 1), (5, 2)}
    # same through the kernel: inject the frame with p=0 so the volume shows the noiseless syndrome
    env = VecSurfaceCodeEnv(5, 0.0, 0.0, "X", False, 5, REF.min_weight(5, "X"), n_envs=1, auto_reset=False)
    w = torch.zeros((env.state_words, env.state_stride), dtype=torch.int64)
    w[0, 0] = 1 << (4 * 6 + 1)
    env.set_state_words(w)
    obs, rew, done, info = env.step(torch.tensor([env.identity_index], dtype=torch.int32, device="cuda"))
    layer = obs[0, 0].cpu().numpy()
    assert {(int(a), int(b)) for a, b in zip(*np.nonzero(layer[::2, ::2]))} == {(4, 1), (5, 2)}
    assert float(rew[0]) == 0.0 and int(info["lifetime"][0]) == 5
    # legal moves = qubits touching those two plaquettes (+ identity)
    assert env.legal_actions(0) == {15, 16, 20, 21, 22, 25}
    env.close()


def test_c3_exact_parameters_bit_exact_against_the_oracle():
    """BASELINE config C3 at its exact parameters -- d=5 depolarising, p_phys = p_meas = 0.007, volume depth 5, 16 384 lattices, the
    SHIPPED referee (the reference's nn_d5_DP_p5 as a table) -- bit for bit against the oracle: explicit-action steps (random-legal picks
    with 10 % arbitrary, possibly illegal or repeated actions), then a rollout under the built-in policy, every output of every step."""
    import ctypes as C
    import torch
    from deepq_decoding_b200 import _lib, referee as R
    n = 16384
    env, o = make_pair(5, "DP", False, 5, 0.007, n, seed=2026, base=0, referee=R.shipped(5, "DP"))
    obs = env.reset().cpu().numpy()
    oobs, olegal = o.reset()
    assert np.array_equal(obs, oobs)
    rng = np.random.default_rng(7)
    ndone = 0
    for t in range(12):
        acts = o.random_legal_actions(olegal, t)
        arb = rng.random(n) < 0.1
        acts[arb] = rng.integers(0, o.A, size=int(arb.sum()))
        obs, rew, done, info = env.step(torch.from_numpy(acts).cuda())
        oobs, orew, odone, olife, olegal = o.step(acts, auto_reset=True)
        assert np.array_equal(obs.cpu().numpy(), oobs), "obs t=%d" % t
        assert np.array_equal(rew.cpu().numpy(), orew) and np.array_equal(done.cpu().numpy(), odone.astype(bool)), "reward / done t=%d" % t
        assert np.array_equal(info["lifetime"].cpu().numpy(), olife) and np.array_equal(info["legal_mask"].cpu().numpy().view(np.uint64), olegal)
        ndone += int(odone.sum())
    L = _lib.lib()
    p = lambda x: C.c_void_p(x.data_ptr())
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    S, slots = 40, 3
    ring = torch.zeros((slots,) + tuple(env.obs.shape), dtype=torch.uint8, device="cuda")
    rew = torch.zeros((S, n), dtype=torch.float32, device="cuda"); done = torch.zeros((S, n), dtype=torch.uint8, device="cuda")
    life = torch.zeros((S, n), dtype=torch.int32, device="cuda"); acts_out = torch.zeros((S, n), dtype=torch.int32, device="cuda")
    legal = torch.zeros((S, n, env.mask_words), dtype=torch.int64, device="cuda")
    _lib.check(L.dq_policy_seek(env._h, 100, st))
    _lib.check(L.dq_env_rollout_random(env._h, S, p(ring), slots, 1, p(rew), p(done), p(life), p(legal), p(acts_out), 1, st))
    torch.cuda.synchronize()
    for s_ in range(S):
        oa = o.random_legal_actions(olegal, 100 + s_)
        oobs, orew, odone, olife, olegal = o.step(oa, auto_reset=True)
        assert np.array_equal(acts_out[s_].cpu().numpy(), oa), "pick s=%d" % s_
        assert np.array_equal(rew[s_].cpu().numpy(), orew) and np.array_equal(done[s_].cpu().numpy(), odone) and np.array_equal(life[s_].cpu().numpy(), olife)
        assert np.array_equal(legal[s_].cpu().numpy().view(np.uint64), olegal)
        ndone += int(odone.sum())
        if s_ >= S - slots:
            assert np.array_equal(ring[(1 + s_) % slots].cpu().numpy(), oobs), "ring obs s=%d" % s_
    assert ndone > 1000, "the run must contain finished episodes (referee decisions)"
    compare_state(env, o, rng.integers(0, n, size=16))
    env.close()


def test_attempt_cap_accepts_a_trivial_volume():
    """Documented deviation (dq_env_set_max_attempts): at p = 0 the reference would redraw the all-trivial volume for ever."""
    from deepq_decoding_b200 import _lib
    env, o = make_pair(3, "DP", False, 3, 0.0, 40, seed=4)
    _lib.check(_lib.lib().dq_env_set_max_attempts(env._h, 3)); o.set_max_attempts(3)
    obs = env.reset().cpu().numpy()
    oobs, olegal = o.reset()
    assert np.array_equal(obs, oobs) and not obs[:, :, ::2, ::2].any()
    for t in range(3):
        acts = o.random_legal_actions(olegal, t)
        import torch
        obs, rew, done, info = env.step(torch.from_numpy(acts).cuda())
        oobs, orew, odone, olife, olegal = o.step(acts, auto_reset=True)
        assert np.array_equal(obs.cpu().numpy(), oobs) and np.array_equal(info["lifetime"].cpu().numpy(), olife)
        assert (olife == (t + 2) * 9).all()
    env.close()


def test_full_size_properties():
    """BASELINE config 3 shape (d=5 DP p=0.007, 16384 lattices): size-independent invariants."""
    import torch
    from deepq_decoding_b200.envs import VecSurfaceCodeEnv
    env = VecSurfaceCodeEnv(5, 0.007, 0.007, "DP", False, 5, None, n_envs=16384, seed=1)
    obs = env.reset()
    marker = obs[0, 0].clone(); marker[::2, ::2] = 0
    heavy_frac = []
    for t in range(50):
        acts = env.random_legal_actions(t).clone()
        legal = env.legal_mask.clone()
        before = (env.get_state_words()[2, :16384] & 0xFFFFFFFF).to(torch.int32)     # stored lifetime
        # the policy only ever picks legal actions
        assert bool(((legal[:, 0] >> acts.long()) & 1).all())
        obs, rew, done, info = env.step(acts)
        # cells are 0/1 and every syndrome layer carries the constant marker pattern
        assert int(obs.max()) <= 1
        m = obs[:, :5].clone(); m[:, :, ::2, ::2] = 0
        assert bool((m == marker).all())
        # lifetime moves in multiples of volume_depth, only on identity / repeated actions
        dl = info["lifetime"] - before
        assert bool((dl % 5 == 0).all()) and bool((dl >= 0).all())
        heavy = (acts == 50) | (dl > 0)
        assert bool(((dl > 0) == heavy).all())
        heavy_frac.append(float((dl > 0).float().mean()))
        # a fresh volume is never trivial and clears the action layers; identity is always legal
        fresh = (dl > 0) | done
        assert bool((obs[fresh][:, :5, ::2, ::2].sum(dim=(1, 2, 3)) > 0).all())
        assert int(obs[fresh][:, 5:].sum()) == 0
        assert bool(((info["legal_mask"][:, 0] >> 50) & 1).all())
        # reward 1 implies not done
        assert not bool(((rew == 1.0) & done).any())
    assert 0.05 < np.mean(heavy_frac) < 0.25          # 11.7 % heavy under random-legal
    env.close()


def test_full_size_c5_rollout_properties():
    """BASELINE config 5 per-GPU shape (d=7 DP p=0.011, 8192 lattices, min-weight referee): a 40-step rollout launch equals 40
    single-step launches on every lattice, and the size-independent invariants hold on its outputs."""
    import ctypes as C
    import torch
    from deepq_decoding_b200 import _lib
    from deepq_decoding_b200.envs import VecSurfaceCodeEnv
    n, S, slots = 8192, 40, 3
    a = VecSurfaceCodeEnv(7, 0.011, 0.011, "DP", False, 7, None, n_envs=n, seed=9)
    b = VecSurfaceCodeEnv(7, 0.011, 0.011, "DP", False, 7, None, n_envs=n, seed=9)
    L = _lib.lib()
    p = lambda x: C.c_void_p(x.data_ptr())
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    obs0 = a.reset().clone()
    assert torch.equal(obs0, b.reset())
    marker = obs0[0, 0].clone(); marker[::2, ::2] = 0
    legal_before = a.legal_mask.clone()
    ring = torch.zeros((slots,) + tuple(a.obs.shape), dtype=torch.uint8, device="cuda")
    out = a.rollout_random(S, ring, keep=("reward", "done", "lifetime", "actions", "legal"))
    ident = a.num_actions - 1
    picks = torch.zeros(n, dtype=torch.int32, device="cuda")
    for s in range(S):
        _lib.check(L.dq_env_step_random(b._h, p(b.obs), p(b.reward), p(b.done), p(b.lifetime), p(b.legal_mask), p(picks), 1, st))
        assert torch.equal(out["actions"][s], picks) and torch.equal(out["done"][s], b.done) and torch.equal(out["lifetime"][s], b.lifetime), s
        # the pick is a legal action of the mask the previous step published; identity is always legal
        act = out["actions"][s].long()
        word = torch.gather(legal_before, 1, (act >> 6)[:, None])[:, 0]
        assert bool(((word >> (act & 63)) & 1).all()), s
        legal_before = out["legal"][s]
        assert bool(((legal_before[:, ident >> 6] >> (ident & 63)) & 1).all())
        assert bool((out["lifetime"][s] % 7 == 0).all())
        assert not bool(((out["reward"][s] == 1.0) & (out["done"][s] != 0)).any())
    assert torch.equal(a.get_state_words(), b.get_state_words())
    last = ring[(S - 1) % slots]
    assert torch.equal(last, b.obs) and int(last.max()) <= 1
    m = last[:, :7].clone(); m[:, :, ::2, ::2] = 0
    assert bool((m == marker).all())
    a.close(); b.close()


def test_errors_are_loud():
    from deepq_decoding_b200 import _lib
    from deepq_decoding_b200.envs import VecSurfaceCodeEnv
    with pytest.raises(Exception):
        VecSurfaceCodeEnv(4, 0.01, 0.01, "DP", False, 3, None, n_envs=4)          # even d (FL:38-39)
    with pytest.raises(ValueError):
        VecSurfaceCodeEnv(5, 0.01, 0.01, "IIDXZ", False, 3, None, n_envs=4)
    with pytest.raises(_lib.DQError):
        VecSurfaceCodeEnv(9, 0.01, 0.01, "DP", False, 3, REF.RefereeLUT(9, "DP", 1, np.zeros(4, np.uint8)), n_envs=4)
    with pytest.raises(_lib.DQError):
        VecSurfaceCodeEnv(5, 0.01, 0.01, "DP", False, 9, None, n_envs=4)          # volume_depth > 8


def test_reference_named_adapter_matches_oracle():
    """The N=1 class with the reference's name, ctor and attribute set, step by step against the oracle."""
    from deepq_decoding_b200.envs import Surface_Code_Environment_Multi_Decoding_Cycles
    rng = np.random.default_rng(8)
    ref = random_referee(rng, 5, "DP")
    env = Surface_Code_Environment_Multi_Decoding_Cycles(d=5, p_phys=0.03, p_meas=0.03, error_model="DP", use_Y=False,
                                                         volume_depth=5, static_decoder=ref, seed=4, env_id=3)
    o = O.OracleVecEnv(5, "DP", False, 5, 0.03, 0.03, 1, 4, 3)
    o.set_referee(ref.mode, ref.lut_a, ref.lut_b)
    assert env.observation_space.shape == (7, 11, 11) and env.action_space.n == 51
    assert env.num_actions == 51 and env.identity_index == 50 and env.n_action_layers == 2 and env.multi_cycle
    board = env.reset(); oobs, olegal = o.reset()
    assert board.dtype.kind == "i" and np.array_equal(board, oobs[0])
    ndone = 0
    for t in range(300):
        legal = sorted(env.legal_actions)
        assert legal == [a for a in range(51) if (int(olegal[0, 0]) >> a) & 1]
        a = legal[int(rng.integers(0, len(legal)))] if rng.random() > 0.2 else int(rng.integers(0, 51))
        board, reward, done, info = env.step(a)
        oobs, orew, odone, olife, olegal = o.step(np.array([a], np.int32), auto_reset=False)
        st = o.get_env(0)
        assert np.array_equal(board, oobs[0]) and board is env.board_state and info == {}
        assert reward == float(orew[0]) and done == bool(odone[0]) and env.done == done and env.lifetime == int(olife[0])
        assert np.array_equal(env.hidden_state, st["hidden"]) and np.array_equal(env.current_true_syndrome, st["true_syndrome"])
        if done:
            ndone += 1
            board = env.reset(); oobs, olegal = o.reset()
            assert np.array_equal(board, oobs[0]) and env.lifetime == o.get_env(0)["lifetime"] and not env.done
    assert ndone > 0
    text = env.render()
    assert isinstance(text, str) and "lifetime" in text
    # constant-layout helpers kept for API parity (notebook 3 builds its input volume with them)
    syn = np.zeros((6, 6), int); syn[4, 1] = syn[5, 2] = 1
    pad = env.padding_syndrome(syn)
    assert pad.shape == (11, 11) and pad[8, 2] == 1 and pad[10, 4] == 1 and pad[0, 1] == 1 and pad[1, 3] == 1 and pad[1, 1] == 0
    assert env.padding_actions([0] * 21 + [1] + [0] * 3)[9, 3] == 1
    env.close()


def test_error_rate_sweep_driver():
    """Single_Point_Training_Script.py:187-222 on the shipped d5_dp/0.007 agent, coarse grid."""
    import sys, os
    sys.path.insert(0, os.path.dirname(__file__))
    from qnet_util import REF_CC, REF_FF, golden_weights
    from deepq_decoding_b200 import agents as A
    from deepq_decoding_b200.envs import VecSurfaceCodeEnv
    from deepq_decoding_b200.evaluate import test_sweep
    conv, dense = golden_weights("dp")
    env = VecSurfaceCodeEnv(5, 0.007, 0.007, "DP", False, 5, None, n_envs=4096, seed=77)
    dqn = A.DQNAgent(model=A.build_convolutional_nn(REF_CC, REF_FF, env.observation_space.shape, env.num_actions), nb_actions=env.num_actions,
                     memory=A.SequentialMemory(limit=100), policy=A.GreedyQPolicy(masked_greedy=True),
                     test_policy=A.GreedyQPolicy(masked_greedy=True), enable_dueling_network=True)
    dqn.compile(A.Adam(lr=1e-5), max_envs=4096)
    dqn.model.set_keras_weights(conv, dense)
    all_results, detailed, stats = test_sweep(dqn, env, nb_test_episodes=4096, num_to_test=4, step=0.005)
    keys = list(all_results)
    assert keys[0] == "0.005" and keys[1] == "0.01"
    assert abs(all_results["0.005"] - 698.57) < 5 * max(stats["0.005"]["standard_error"], 1.0)
    assert all_results["0.005"] > 200 > all_results["0.01"]
    assert all_results[keys[-1]] < 1.0 / float(keys[-1]) or len(keys) == 4                       # stopped below 1/p
    assert detailed["0.005"][-1] == pytest.approx(all_results["0.005"])
    env.close()
