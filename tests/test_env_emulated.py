"""CPU check of the CUDA environment kernel's SOURCE: `dq_env.cu` compiled by g++ under `tests/host/cuda_emu.h`
(every CUDA thread a fiber, warp primitives and barriers as rendez-vous) and driven through the same C ABI,
compared bit for bit with the oracle.

This does not replace the `-m gpu` parity tests (those run the sm_100a build on a B200); it is what lets a
kernel change be validated in this GPU-less container before GPU minutes are spent on it.  Cases mirror
tests/test_env_gpu.py at sizes the fibers finish in seconds.
"""
import numpy as np
import pytest

from oracle import oracle as O
import emu_env as E

JOINT, SPLIT = 0, 1


def random_luts(rng, d, model):
    ns = d * d - 1
    if model == "X" or d == 7:
        la = rng.integers(0, 256, size=(1 << (ns // 2)) // 4 + 1, dtype=np.uint8) & 0x55
        lb = rng.integers(0, 256, size=(1 << (ns // 2)) // 4 + 1, dtype=np.uint8) & 0x55
        return SPLIT, la, (lb if model == "DP" else None)
    return JOINT, rng.integers(0, 256, size=(1 << ns) // 4 + 1, dtype=np.uint8), None


def make_pair(d, model, use_Y, vd, p, n, seed, base=0):
    rng = np.random.default_rng(1000 * d + vd)
    mode, la, lb = random_luts(rng, d, model)
    env = E.EmuVecEnv(d, model, use_Y, vd, p, p, n, seed, base)
    o = O.OracleVecEnv(d, model, use_Y, vd, p, p, n, seed, base)
    env.set_referee(mode, la, lb)
    o.set_referee(mode, la, lb)
    return env, o


def grid_to_pauli(xw, zw, d):
    g = d + 1
    out = np.zeros((d, d), np.int8)
    for r in range(d):
        for c in range(d):
            x, z = (int(xw) >> (r * g + c)) & 1, (int(zw) >> (r * g + c)) & 1
            out[r, c] = 2 if (x and z) else (1 if x else (3 if z else 0))
    return out


def compare_state(env, o, idx):
    w = env.state_words()
    for i in idx:
        ost = o.get_env(int(i))
        meta = int(w[E.ROW_META, i])
        assert np.array_equal(grid_to_pauli(w[E.ROW_XB, i], w[E.ROW_ZB, i], env.d), ost["hidden"])
        assert (meta & 0xFFFFFFFF) == ost["lifetime"] and ((meta >> 32) & 0x7FFFFFFF) == ost["attempts"]
        assert bool(meta >> 63) == ost["done"]


CASES = [(3, "X", False, 3, 0.05, 1),
         (3, "X", False, 3, 0.05, 100),
         (5, "X", False, 5, 0.02, 45),        # ragged tail CTA
         (5, "DP", False, 5, 0.02, 64),
         (5, "DP", True, 3, 0.03, 33),
         (7, "DP", False, 7, 0.011, 40),
         (7, "DP", True, 8, 0.02, 17),        # deepest volume, three mask words
         (3, "DP", True, 2, 0.08, 48)]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "d%d_%s_y%d_vd%d_n%d" % (c[0], c[1], c[2], c[3], c[5]))
def test_emulated_kernel_trajectory_parity(case):
    d, model, use_Y, vd, p, n = case
    env, o = make_pair(d, model, use_Y, vd, p, n, seed=42, base=5)
    rng = np.random.default_rng(d + n)
    obs, legal = env.reset()
    oobs, olegal = o.reset()
    assert np.array_equal(obs, oobs) and np.array_equal(legal, olegal)
    ndone = 0
    for t in range(60):
        acts = env.random_legal_actions(legal, t)
        assert np.array_equal(acts, o.random_legal_actions(olegal, t)), "random-legal policy"
        arb = rng.random(n) < 0.2
        acts[arb] = rng.integers(-1, o.A + 1, size=int(arb.sum()))      # includes out-of-range -> identity
        oacts = np.where((acts < 0) | (acts >= o.A), o.A - 1, acts).astype(np.int32)
        obs, rew, done, life, legal = env.step(acts)
        oobs, orew, odone, olife, olegal = o.step(oacts, auto_reset=True)
        assert np.array_equal(obs, oobs), "obs t=%d" % t
        assert np.array_equal(rew, orew) and np.array_equal(done, odone), "reward / done t=%d" % t
        assert np.array_equal(life, olife), "lifetime t=%d" % t
        assert np.array_equal(legal, olegal), "legal t=%d" % t
        ndone += int(odone.sum())
        if t % 20 == 0:
            compare_state(env, o, rng.integers(0, n, size=min(n, 4)))
    if n >= 40:
        assert ndone > 0, "the trajectory should contain finished episodes"


@pytest.mark.parametrize("d,model,n,auto_reset", [(5, "DP", 50, True), (3, "X", 37, True), (7, "DP", 20, True), (5, "X", 30, False)])
def test_emulated_rollout_equals_single_steps_and_oracle(d, model, n, auto_reset):
    """dq_env_rollout_random (one launch, many steps, observation ring) == dq_env_step_random launches == oracle."""
    vd, p, steps, slots = d, 0.03, 24, 5
    a, o = make_pair(d, model, False, vd, p, n, seed=7, base=3)
    b, _ = make_pair(d, model, False, vd, p, n, seed=7, base=3)
    _, olegal = o.reset()
    a.reset(); b.reset()
    a.policy_seek(11); b.policy_seek(11)
    ring, rew, done, life, legal, acts = a.rollout_random(steps, slots, first_slot=2, auto_reset=auto_reset)
    for s in range(steps):
        oacts = o.random_legal_actions(olegal, 11 + s)
        oobs, orew, odone, olife, olegal = o.step(oacts, auto_reset=auto_reset)
        sobs, srew, sdone, slife, slegal, sacts = b.step_random(auto_reset=auto_reset)
        assert np.array_equal(acts[s], oacts) and np.array_equal(sacts, oacts), "actions s=%d" % s
        assert np.array_equal(rew[s], orew) and np.array_equal(srew, orew)
        assert np.array_equal(done[s], odone) and np.array_equal(sdone, odone)
        assert np.array_equal(life[s], olife) and np.array_equal(slife, olife)
        assert np.array_equal(legal[s], olegal) and np.array_equal(slegal, olegal)
        assert np.array_equal(sobs, oobs), "single-step obs s=%d" % s
        if s >= steps - slots:                         # the ring keeps the last `slots` observations
            assert np.array_equal(ring[(2 + s) % slots], oobs), "ring obs s=%d" % s
    assert np.array_equal(a.state_words(), b.state_words())
    # the device-side step counter advanced by the number of steps on both
    nxt_a = a.step_random(auto_reset=auto_reset)[5]
    assert np.array_equal(nxt_a, o.random_legal_actions(olegal, 11 + steps))


def test_emulated_host_entry_points_and_unaligned_observation_buffers():
    """dq_env_step_host, and observation pointers at 8-byte / odd alignment (the 64-bit and byte store paths of phase D)."""
    d, model, n = 5, "DP", 21
    env, o = make_pair(d, model, False, 5, 0.03, n, seed=3)
    obs, legal = env.reset()
    oobs, olegal = o.reset()
    assert np.array_equal(obs, oobs)
    for t in range(6):
        acts = o.random_legal_actions(olegal, t)
        got = env.step_host(acts)
        want = o.step(acts, auto_reset=True)
        for g, w in zip(got, want):
            assert np.array_equal(g, w)
        olegal = want[4]
    import ctypes as C
    for off in (8, 3):
        acts = o.random_legal_actions(olegal, 100 + off)
        nbytes = n * env.Cn * env.H * env.H
        raw = np.zeros(nbytes + 64, np.uint8)
        base = raw.ctypes.data
        start = (-base) % 16 + off
        view = raw[start:start + nbytes]
        reward, done, life, legal2, _ = env._outs()
        env._check(env.L.dq_env_step(env.h, E._p(acts), C.c_void_p(base + start), E._p(reward), E._p(done), E._p(life),
                                     E._p(legal2), 1, None))
        want = o.step(acts, auto_reset=True)
        assert np.array_equal(view.reshape(want[0].shape), want[0]), "observation at alignment %d" % off
        assert not raw[:start].any() and not raw[start + nbytes:].any(), "bytes outside the buffer were written"
        olegal = want[4]


def test_emulated_host_calls_plain_and_threaded_expansion():
    """dq_env_reset_host / dq_env_step_host return the same byte observations whether the library moves bit-packed rows and expands
    them with its host threads (the default; any thread count; AVX2 or the table walk) or lets the kernel write bytes and copies all
    of them back (DQ_HOST_EXPAND=0).  The switches are read once per process, hence the child processes."""
    import os, subprocess, sys
    code = (
        "import sys, numpy as np\n"
        "sys.path.insert(0, %r)\n"
        "import emu_env as E\n"
        "from test_env_emulated import make_pair\n"
        "for d, model, use_Y, vd, n in ((5, 'DP', False, 5, 300), (7, 'DP', True, 4, 19), (3, 'X', False, 3, 257)):\n"
        "    env, o = make_pair(d, model, use_Y, vd, 0.03, n, seed=3)\n"
        "    assert env._info(10) == int(sys.argv[1])\n"
        "    obs, legal = env.reset_host()\n"
        "    oobs, olegal = o.reset()\n"
        "    assert np.array_equal(obs, oobs) and np.array_equal(legal, olegal)\n"
        "    for t in range(12):\n"
        "        acts = o.random_legal_actions(olegal, t)\n"
        "        got, want = env.step_host(acts), o.step(acts, auto_reset=True)\n"
        "        assert all(np.array_equal(g, w) for g, w in zip(got, want)), (d, t)\n"
        "        olegal = want[4]\n"
        "print('ok')\n"
    ) % os.path.dirname(os.path.abspath(__file__))
    # (the small outputs and the actions: straight into / out of the caller's buffers -- the emulation reports every buffer as pinned --,
    #  through the library's pinned staging block, or by DMA copies)
    for expand, threads, extra in (("1", "1", {}), ("1", "5", {}), ("1", "3", {"DQ_HOST_NO_AVX2": "1"}), ("0", "2", {}),
                                   ("1", "2", {"DQ_HOST_DIRECT": "0"}), ("1", "2", {"DQ_HOST_ZEROCOPY": "0"}), ("0", "2", {"DQ_HOST_ZEROCOPY": "0"})):
        out = subprocess.run([sys.executable, "-c", code, expand], env=dict(os.environ, DQ_HOST_EXPAND=expand, DQ_HOST_THREADS=threads, **extra),
                             capture_output=True, text=True, timeout=600)
        assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]


def test_emulated_split_host_call_and_host_policy():
    """dq_env_step_host_begin / _end on two handles driven alternately (the overlap pattern of the e2e benchmark) and the host-side
    random-legal policy: same picks as the device policy and the oracle, same outputs as the one-piece call."""
    d, model, n = 5, "DP", 4100                       # >= 4096 lattices: the bitmap rows cross in several lattice ranges
    a, oa = make_pair(d, model, False, 5, 0.03, n, seed=8, base=0)
    b, ob = make_pair(d, model, False, 5, 0.03, n, seed=8, base=n)
    la, lb = oa.reset()[1], ob.reset()[1]
    assert np.array_equal(a.reset_host()[1], la) and np.array_equal(b.reset_host()[1], lb)
    for t in range(4):
        acts_a, acts_b = a.random_legal_actions_host(la, t), b.random_legal_actions_host(lb, t)
        assert np.array_equal(acts_a, oa.random_legal_actions(la, t)) and np.array_equal(acts_b, ob.random_legal_actions(lb, t))
        assert np.array_equal(acts_a, a.random_legal_actions(la, t)), "host policy == device policy"
        a.step_host_begin(acts_a)
        b.step_host_begin(acts_b)
        got_a = a.step_host_end()
        want_a = oa.step(acts_a, auto_reset=True)
        got_b = b.step_host_end()
        want_b = ob.step(acts_b, auto_reset=True)
        assert all(np.array_equal(g, w) for g, w in zip(got_a, want_a)) and all(np.array_equal(g, w) for g, w in zip(got_b, want_b)), t
        la, lb = want_a[4], want_b[4]
    with pytest.raises(RuntimeError):
        a.step_host_end()                              # nothing in flight
    a.step_host_begin(oa.random_legal_actions(la, 9))
    with pytest.raises(RuntimeError):
        a.step_host_begin(oa.random_legal_actions(la, 9))      # one call per handle at a time
    a.step_host_end()


def test_emulator_reports_deadlock_free_run_of_every_geometry():
    """Reset alone (RESET=true instantiation) on every supported distance, single lattice and ragged tile."""
    for d in (3, 5, 7):
        for n in (1, 19):
            env, o = make_pair(d, "DP", False, d, 0.05, n, seed=1)
            obs, legal = env.reset()
            oobs, olegal = o.reset()
            assert np.array_equal(obs, oobs) and np.array_equal(legal, olegal)


@pytest.mark.parametrize("p_phys,p_meas", [(0.01, 0.08), (0.08, 0.01), (0.0, 0.05), (0.05, 0.0)])
def test_emulated_distinct_measurement_rate(p_phys, p_meas):
    """p_meas != p_phys: data-qubit and measurement draws use different thresholds (the screen uses the larger one)."""
    d, model, n = 5, "DP", 40
    env, o = make_pair(d, model, False, 4, 0.03, n, seed=11)
    env.set_noise(p_phys, p_meas); o.set_noise(p_phys, p_meas)
    obs, legal = env.reset()
    oobs, olegal = o.reset()
    assert np.array_equal(obs, oobs) and np.array_equal(legal, olegal)
    for t in range(25):
        acts = o.random_legal_actions(olegal, t)
        got = env.step(acts)
        want = o.step(acts, auto_reset=True)
        for k, (g, w) in enumerate(zip(got, want)):
            assert np.array_equal(g, w), "output %d at t=%d" % (k, t)
        olegal = want[4]


VARIANTS = [("w1g1", ["-DDQ_WRITERS=1", "-DDQ_GENS=1"]),
            ("w2g3", ["-DDQ_WRITERS=2", "-DDQ_GENS=3"]),
            ("w2g5", ["-DDQ_WRITERS=2", "-DDQ_GENS=5"]),
            ("w4g3", ["-DDQ_WRITERS=4", "-DDQ_GENS=3"]),
            ("w3g6", ["-DDQ_WRITERS=3", "-DDQ_GENS=6"]),
            ("q2", ["-DDQ_QDEPTH=2"]),
            ("q8", ["-DDQ_QDEPTH=8"])]


@pytest.mark.parametrize("name,flags", VARIANTS, ids=[v[0] for v in VARIANTS])
def test_emulated_build_variants(name, flags):
    """The tuning builds of tools/build_variants.sh (writer / generator warps per CTA, depth of the volume queues) are the same function:
    reset, a single-step launch, a long and a short rollout, then the packed state, all against the oracle."""
    import os
    path = E.build(flags, out=os.path.join(E.HERE, "host", "libdq_env_emu_%s.so" % name))
    for d, model, n, p, ar in [(5, "DP", 50, 0.03, True), (3, "X", 37, 0.05, True), (7, "DP", 20, 0.02, True),
                               (5, "X", 30, 0.03, False), (5, "DP", 16, 0.007, True)]:
        mode, la, lb = random_luts(np.random.default_rng(d), d, model)
        a = E.EmuVecEnv(d, model, False, d, p, p, n, 7, 3, lib_path=path)
        o = O.OracleVecEnv(d, model, False, d, p, p, n, 7, 3)
        a.set_referee(mode, la, lb); o.set_referee(mode, la, lb)
        obs, _ = a.reset()
        oobs, olegal = o.reset()
        assert np.array_equal(obs, oobs)
        a.policy_seek(0)
        t = 0
        for S in (1, 30, 7):
            if S == 1:
                ring, (rew, done, life, legal, acts) = None, [x[None] for x in a.step_random(ar)[1:]]
            else:
                ring, rew, done, life, legal, acts = a.rollout_random(S, 4, 0, ar)
            for s in range(S):
                oa = o.random_legal_actions(olegal, t)
                oobs, orew, odone, olife, olegal = o.step(oa, auto_reset=ar)
                t += 1
                assert np.array_equal(acts[s], oa) and np.array_equal(rew[s], orew) and np.array_equal(done[s], odone)
                assert np.array_equal(life[s], olife) and np.array_equal(legal[s], olegal)
                if ring is not None and s >= S - 4:
                    assert np.array_equal(ring[s % 4], oobs)
        compare_state(a, o, range(min(n, 8)))


def test_emulated_noise_change_discards_queued_volumes():
    """env.p_phys / env.p_meas are assignable between steps (SPTS:200-201): volume attempts queued ahead under the old rates must not
    be served afterwards -- single steps and a rollout either side of two changes, against the oracle."""
    d, model, n = 5, "DP", 40
    env, o = make_pair(d, model, False, 5, 0.02, n, seed=21)
    _, olegal = o.reset()
    env.reset()
    t = 0
    for pp, pm in ((0.05, 0.01), (0.004, 0.03)):
        for _ in range(5):
            acts = o.random_legal_actions(olegal, t)
            got, want = env.step(acts), o.step(acts, auto_reset=True)
            assert all(np.array_equal(g, w) for g, w in zip(got, want)), "before the change, t=%d" % t
            olegal = want[4]; t += 1
        env.set_noise(pp, pm); o.set_noise(pp, pm)
        env.policy_seek(t)
        ring, rew, done, life, legal, acts = env.rollout_random(9, 3, 0, True)
        for s in range(9):
            oa = o.random_legal_actions(olegal, t)
            oobs, orew, odone, olife, olegal = o.step(oa, auto_reset=True); t += 1
            assert np.array_equal(acts[s], oa) and np.array_equal(rew[s], orew) and np.array_equal(done[s], odone), "after the change, s=%d" % s
            assert np.array_equal(life[s], olife) and np.array_equal(legal[s], olegal)
            if s >= 6:
                assert np.array_equal(ring[s % 3], oobs)
    compare_state(env, o, range(8))


def test_emulated_injected_state_revalidates_queues():
    """dq_env_set_state (checkpoint restore / parity injection): a handle that receives another trajectory's state continues that
    trajectory bit for bit, whatever volume attempts its own queues held (their fill counters no longer match the attempt counters)."""
    d, model, n = 5, "DP", 37
    a, o = make_pair(d, model, False, 5, 0.03, n, seed=13, base=2)
    b, _ = make_pair(d, model, False, 5, 0.03, n, seed=13, base=2)
    _, olegal = o.reset()
    a.reset(); b.reset()
    t = 0
    for _ in range(7):                       # only `a` (and the oracle) advance: b's queues stay at the reset position
        acts = o.random_legal_actions(olegal, t)
        a.step(acts); olegal = o.step(acts, auto_reset=True)[4]; t += 1
    w = a.state_words()
    b._check(b.L.dq_env_set_state(b.h, E._p(w), None))
    for _ in range(8):
        acts = o.random_legal_actions(olegal, t)
        got, want = b.step(acts), o.step(acts, auto_reset=True)
        assert all(np.array_equal(g, x) for g, x in zip(got, want)), "restored handle, t=%d" % t
        olegal = want[4]; t += 1
    # and back in time: the first handle receives an EARLIER state than the one its queues were filled for
    c, o2 = make_pair(d, model, False, 5, 0.03, n, seed=13, base=2)
    _, ol2 = o2.reset(); c.reset()
    w0 = c.state_words().copy()
    b._check(b.L.dq_env_set_state(b.h, E._p(w0), None))
    for t2 in range(6):
        acts = o2.random_legal_actions(ol2, t2)
        got, want = b.step(acts), o2.step(acts, auto_reset=True)
        assert all(np.array_equal(g, x) for g, x in zip(got, want)), "rewound handle, t=%d" % t2
        ol2 = want[4]


@pytest.mark.parametrize("variant", [None, "q2"])
def test_emulated_many_trivial_attempts_in_one_step(variant):
    """Low noise on a small code: most volume attempts are all-trivial, so one step pops many more attempts than a queue holds --
    the physics warp must be able to wake idle generator warps from inside a step (a doorbell rung only after the step deadlocks)."""
    import os
    path = None if variant is None else E.build(dict(VARIANTS)[variant], out=os.path.join(E.HERE, "host", "libdq_env_emu_%s.so" % variant))
    for d, model, vd, p in ((3, "X", 1, 0.004), (3, "DP", 2, 0.002), (5, "X", 2, 0.0015)):
        n = 45
        mode, la, lb = random_luts(np.random.default_rng(d), d, model)
        env = E.EmuVecEnv(d, model, False, vd, p, p, n, 77, 0, lib_path=path)
        o = O.OracleVecEnv(d, model, False, vd, p, p, n, 77, 0)
        env.set_referee(mode, la, lb); o.set_referee(mode, la, lb)
        obs, legal = env.reset()
        oobs, olegal = o.reset()
        assert np.array_equal(obs, oobs) and np.array_equal(legal, olegal)
        attempts0 = sum(o.get_env(i)["attempts"] for i in range(n))
        assert attempts0 > 4 * n, "the reset alone should need several attempts per lattice"
        for t in range(6):
            acts = o.random_legal_actions(olegal, t)
            got, want = env.step(acts), o.step(acts, auto_reset=True)
            assert all(np.array_equal(g, w) for g, w in zip(got, want)), (d, t)
            olegal = want[4]
        env.policy_seek(6)
        ring, rew, done, life, legal2, acts = env.rollout_random(10, 2, 0, True)
        for s_ in range(10):
            oa = o.random_legal_actions(olegal, 6 + s_)
            oobs, orew, odone, olife, olegal = o.step(oa, auto_reset=True)
            assert np.array_equal(acts[s_], oa) and np.array_equal(life[s_], olife) and np.array_equal(legal2[s_], olegal)
        assert np.array_equal(ring[1], oobs)
        compare_state(env, o, range(6))


@pytest.mark.parametrize("cap", [1, 3])
def test_emulated_attempt_cap_accepts_a_trivial_volume(cap):
    """Documented deviation (include/dq_decoding.h, dq_env_set_max_attempts): at p_phys = p_meas = 0 the reference would redraw the
    all-trivial volume for ever; after `cap` attempts the kernel accepts it.  The oracle restates the same rule when told to."""
    d, model, n = 3, "DP", 9
    env, o = make_pair(d, model, False, 3, 0.0, n, seed=4)
    env.set_max_attempts(cap); o.set_max_attempts(cap)
    obs, legal = env.reset()
    oobs, olegal = o.reset()
    assert np.array_equal(obs, oobs) and np.array_equal(legal, olegal)
    assert not obs[:, :, ::2, ::2].any(), "an all-trivial volume shows no syndrome"
    for t in range(4):
        acts = o.random_legal_actions(olegal, t)          # only the identity is legal: every step draws a volume
        got, want = env.step(acts), o.step(acts, auto_reset=True)
        assert all(np.array_equal(g, w) for g, w in zip(got, want)), "t=%d" % t
        assert (want[3] == (t + 2) * cap * 3).all(), "lifetime counts every attempt's slices"
        olegal = want[4]
    compare_state(env, o, range(n))


@pytest.mark.parametrize("d,model,use_Y,vd,n", [(5, "DP", False, 5, 45), (7, "DP", True, 4, 19), (3, "X", False, 3, 33)])
def test_emulated_packed_host_calls(d, model, use_Y, vd, n):
    """dq_env_reset_host_packed / dq_env_step_host_packed: the bit-packed observation rows, expanded on the host by
    deepq_decoding_b200.envs.unpack_observations, are the oracle's byte observations; every other output as usual."""
    from deepq_decoding_b200.envs import unpack_observations
    env, o = make_pair(d, model, use_Y, vd, 0.03, n, seed=5)
    packed, legal = env.reset_host_packed()
    oobs, olegal = o.reset()
    assert packed.shape == (env.Cn * ((env.H * env.H + 63) // 64), env.stride)
    assert np.array_equal(unpack_observations(packed, n, d, env.Cn), oobs) and np.array_equal(legal, olegal)
    for t in range(30):
        acts = o.random_legal_actions(olegal, t)
        packed, rew, done, life, legal = env.step_host_packed(acts)
        oobs, orew, odone, olife, olegal = o.step(acts, auto_reset=True)
        assert np.array_equal(unpack_observations(packed, n, d, env.Cn), oobs), "packed obs t=%d" % t
        assert np.array_equal(rew, orew) and np.array_equal(done, odone) and np.array_equal(life, olife)
        assert np.array_equal(legal, olegal)
        assert not packed[:, n:].any(), "padding lattices stay empty"


@pytest.mark.parametrize("seed0", [100, 200, 300])
def test_emulated_random_configurations(seed0):
    """Randomised sweep over the whole parameter space of the C ABI (d, noise model, use_Y, volume depth 1..8, lattice count,
    stream id base, p_phys / p_meas from 0 to 0.2, auto-reset on or off): explicit actions (30 % arbitrary, some out of range),
    single-step launches and rollouts of random length into rings of random size, then the packed state -- all against the oracle."""
    lut_cache = {}
    for it in range(6):
        rng = np.random.default_rng(seed0 + it)
        d = int(rng.choice([3, 5, 7])); model = str(rng.choice(["X", "DP"])); use_Y = bool(rng.integers(0, 2)) and model == "DP"
        vd = int(rng.integers(1, 9)); n = int(rng.integers(1, 70)); base = int(rng.integers(0, 1000)); seed = int(rng.integers(0, 2 ** 62))
        pp = float(rng.choice([0.0, 0.002, 0.01, 0.05, 0.2])); pm = float(rng.choice([0.0, 0.002, 0.01, 0.05, 0.2]))
        if pp == 0.0 and pm == 0.0:
            pm = 0.01
        ar = bool(rng.integers(0, 2))
        cfg = dict(d=d, model=model, use_Y=use_Y, vd=vd, n=n, base=base, seed=seed, p_phys=pp, p_meas=pm, auto_reset=ar)
        if (d, model) not in lut_cache:
            lut_cache[(d, model)] = random_luts(np.random.default_rng(d * 7 + len(model)), d, model)
        mode, la, lb = lut_cache[(d, model)]
        env = E.EmuVecEnv(d, model, use_Y, vd, pp, pm, n, seed, base)
        o = O.OracleVecEnv(d, model, use_Y, vd, pp, pm, n, seed, base)
        env.set_referee(mode, la, lb); o.set_referee(mode, la, lb)
        obs, legal = env.reset()
        oobs, olegal = o.reset()
        assert np.array_equal(obs, oobs) and np.array_equal(legal, olegal), cfg
        t = 0
        for _ in range(3):
            if int(rng.integers(0, 3)) == 0:
                for _ in range(6):
                    acts = o.random_legal_actions(olegal, t)
                    arb = rng.random(n) < 0.3
                    acts[arb] = rng.integers(-2, o.A + 2, size=int(arb.sum()))
                    oacts = np.where((acts < 0) | (acts >= o.A), o.A - 1, acts).astype(np.int32)
                    got, want = env.step(acts, auto_reset=ar), o.step(oacts, auto_reset=ar)
                    olegal = want[4]; t += 1
                    assert all(np.array_equal(g, w) for g, w in zip(got, want)), cfg
            else:
                S, slots = int(rng.integers(1, 14)), int(rng.integers(1, 5))
                fs = int(rng.integers(0, slots))
                env.policy_seek(t)
                ring, rew, done, life, legal2, acts = env.rollout_random(S, slots, fs, ar)
                for s in range(S):
                    oa = o.random_legal_actions(olegal, t)
                    oobs, orew, odone, olife, olegal = o.step(oa, auto_reset=ar); t += 1
                    assert np.array_equal(acts[s], oa) and np.array_equal(rew[s], orew) and np.array_equal(done[s], odone), cfg
                    assert np.array_equal(life[s], olife) and np.array_equal(legal2[s], olegal), cfg
                    if s >= S - slots:
                        assert np.array_equal(ring[(fs + s) % slots], oobs), cfg
        compare_state(env, o, range(min(n, 6)))
