#!/usr/bin/env python
"""Benchmark of the hot path: env-steps/s at d=5 depolarising p=0.007 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config C3 of SURVEY.md section 8): d=5, depolarising noise, p_phys = p_meas = 0.007, volume depth 5,
16384 lattices per GPU (weak scaling: C4 = 8 x 16384), shipped referee table, random-legal policy.
One "step" = one vectorised env step of every lattice: the policy kernel picks an action per lattice
from its legal mask, the env kernel advances all lattices and writes a fresh observation (both in one launch:
dq_env_step_random).

  value     all-GPU env-steps/s with everything resident in HBM; the K steps run as rollout launches
            (dq_env_rollout_random, 256 steps of every lattice per launch, bit-identical to single-step
            launches), step s writing its observations into slot s % 16 of a 16-slot ring (222 MB > the
            126 MB L2, so observation writes cannot be absorbed by L2) and row s of the per-step outputs;
            `single_step_launches` repeats the K steps as one launch per step (CUDA graph of 16)
  roofline  env-step kernel only: algorithmic bytes per launch (996 B per lattice-step x
            lattices x steps per launch) over its mean duration, measured with CUDA events around every
            launch of a pass queued behind a device-side delay so the events see back-to-back kernels
  e2e       the same steps through dq_env_step_host: actions come from pinned host memory, every output
            (observations, reward, done, lifetime, legal mask) is copied back to the host each step
  e2e_packed
            the e2e loop with dq_env_step_host_packed: observations cross PCIe as the bit-packed rows the Q-network
            consumes (7.5x fewer bytes); reported beside e2e, which stays the byte-observation number
  cpu_baseline / --impl reference
            the CPU oracle (oracle/dq_oracle.c, a C restatement of the reference env; the Python
            reference itself cannot travel to the GPU box) on all host cores, same workload
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line: keep NCCL's "NCCL version ..." banner (NCCL_DEBUG=VERSION, from the environment or
# /etc/nccl.conf) off it
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"

D, VD, P, MODEL, USE_Y = 5, 5, 0.007, "DP", False
N_PER_GPU = 16384
# other BASELINE.json configs, selectable with --workload for the roofline report (the default, c3, is the one the metric is quoted on)
WORKLOADS = {            # d, vd, p, model, lattices per GPU, algorithmic bytes per lattice-step
    "c1": (3, 3, 0.05, "X", 1, 345),
    "c2": (5, 5, 0.007, "X", 4096, 875),
    "c3": (5, 5, 0.007, "DP", 16384, 996),
    "c5": (7, 7, 0.011, "DP", 8192, 2310),
}
SEED = 2026
RING = 16
ROLL = 256         # env steps per rollout launch on the timed path
METRIC = "env-steps/sec at d=5 depolarising p=0.007"
UNIT = "env-steps/s"
BYTES_PER_STEP = 996          # obs 847 + action 4 + reward 4 + done 1 + lifetime 4 + mask 8 + 2 x 64 state
WORKLOAD = ("C3: d=5 DP p_phys=p_meas=0.007 volume_depth=5 use_Y=False, %d lattices/GPU, random-legal policy, "
            "referee = shipped nn_d5_DP_p5 tabulated" % N_PER_GPU)


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------- CPU arm
def cpu_oracle_run(budget_s, n=N_PER_GPU):
    """Times the oracle on all host cores: random-legal policy + step, like the GPU loop."""
    import numpy as np
    from oracle import oracle as O
    from deepq_decoding_b200 import referee as REF
    try:
        ncores = len(os.sched_getaffinity(0))
    except Exception:
        ncores = os.cpu_count() or 1
    O.lib().dqo_set_num_threads(ncores)          # torchrun exports OMP_NUM_THREADS=1: use every core we are allowed
    o = O.OracleVecEnv(D, MODEL, USE_Y, VD, P, P, n, SEED)
    ref = REF.shipped(D, MODEL)
    o.set_referee(ref.mode, ref.lut_a, ref.lut_b)
    _, legal = o.reset()
    cores = O.lib().dqo_num_threads()
    for t in range(3):
        legal = o.step(o.random_legal_actions(legal, t), auto_reset=True)[4]
    steps, t0 = 0, time.perf_counter()
    while True:
        acts = o.random_legal_actions(legal, 3 + steps)
        legal = o.step(acts, auto_reset=True)[4]
        steps += 1
        el = time.perf_counter() - t0
        if el >= budget_s:
            break
    return dict(value=n * steps / el, unit=UNIT, cores=cores, kind="port",
                sample="%d vectorised steps of %d lattices (%.1f s wall) of the same workload, oracle/dq_oracle.c "
                       "with OpenMP over lattices" % (steps, n, el)), n * steps / el, el / steps


def cpu_serial_dqn_loop(budget_s):
    """The reference's training loop shape on one host core: ONE lattice, per env step one Q-network forward for
    the eps-greedy pick and one double-DQN update on a batch of 32 (keras-rl defaults: train_interval 1, batch 32).  The env is
    the CPU oracle and the network the torch-CPU restatement (oracle/qnet_ref.py) -- both stand-ins for the reference's numpy /
    Keras code, which cannot travel to the GPU box; the authors' own log averages 41.6 steps/s."""
    import copy
    import numpy as np
    import torch
    from oracle import oracle as O, qnet_ref as R
    from deepq_decoding_b200 import referee as REF
    torch.set_num_threads(1)
    o = O.OracleVecEnv(D, MODEL, USE_Y, VD, P, P, 1, SEED + 11)
    ref = REF.shipped(D, MODEL)
    o.set_referee(ref.mode, ref.lut_a, ref.lut_b)
    rng = np.random.default_rng(SEED)
    A_n, C_n, H = o.A, o.Cn, o.H
    conv, dense = R.glorot_uniform_params(rng, C_n, [(64, 3, 2), (32, 2, 1), (32, 2, 1)], [512], A_n, H, dueling=True)
    net = R.TorchQNet(conv, dense, [2, 1, 1], dueling=True)
    target = R.TorchQNet(copy.deepcopy(conv), copy.deepcopy(dense), [2, 1, 1], dueling=True)
    params = net.parameters()
    m = [torch.zeros_like(x) for x in params]
    v = [torch.zeros_like(x) for x in params]
    obs, legal = o.reset()
    mem_s, mem_a, mem_r, mem_t, mem_s1 = [], [], [], [], []
    steps, updates, t0 = 0, 0, time.perf_counter()
    while time.perf_counter() - t0 < budget_s:
        with torch.no_grad():
            q = net.forward(obs)[0].numpy()
        lw = legal[0]
        legal_ids = [a for a in range(A_n) if (int(lw[a >> 6]) >> (a & 63)) & 1]
        a = int(rng.choice(legal_ids)) if rng.random() < 0.1 else int(legal_ids[int(np.argmax(q[legal_ids]))])
        obs1, rew, done, _, legal = o.step(np.array([a], np.int32), auto_reset=True)
        mem_s.append(obs[0].copy()); mem_a.append(a); mem_r.append(float(rew[0])); mem_t.append(float(done[0])); mem_s1.append(obs1[0].copy())
        obs = obs1
        steps += 1
        if len(mem_a) >= 32:
            idx = rng.integers(0, len(mem_a), size=32)
            s0 = np.stack([mem_s[i] for i in idx]); s1 = np.stack([mem_s1[i] for i in idx])
            with torch.no_grad():
                y = R.dqn_targets(net.forward(s1), target.forward(s1), torch.tensor([mem_r[i] for i in idx]),
                                  torch.tensor([mem_t[i] for i in idx]), 0.99)
            loss = R.dqn_loss(net.forward(s0), torch.tensor([mem_a[i] for i in idx]), y)
            grads = torch.autograd.grad(loss, params)
            updates += 1
            R.keras_adam_step(params, grads, m, v, updates, 1e-4)
    el = time.perf_counter() - t0
    return {"env_steps_per_s": steps / el, "updates": updates, "seconds": el, "cores": 1,
            "what": "one lattice, forward + eps-greedy + oracle env step + one batch-32 double-DQN update per step, torch-CPU fp32 "
                    "(the reference's loop shape; its own log: 41.6 steps/s)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t_budget = min(60.0, max(5.0, 0.02 * (args.steps + args.warmup)))
    cb, value, s_per_step = cpu_oracle_run(t_budget)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": s_per_step * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "CPU oracle port of the reference env on the host cores; each step is "
                       "one vectorised step of 16384 lattices; the run is time-bounded, not step-bounded"},
            "cpu_baseline": dict(cb, value=value),
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1] or [r for _, r in self.rows[-3:]]
        sm = sorted(float(r[0]) for r in rows if r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows for i in range(4) if len(r) >= 8 and r[4 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(rows)}


# --------------------------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from deepq_decoding_b200 import _lib
    from deepq_decoding_b200.envs import VecSurfaceCodeEnv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    K, Wm = args.steps, max(args.warmup, 3)
    n = N_PER_GPU
    L = _lib.lib()

    env = VecSurfaceCodeEnv(D, P, P, MODEL, USE_Y, VD, None, n_envs=n, seed=SEED, env_id_base=rank * n, device=dev)
    h = env._h
    ring = torch.zeros((RING,) + tuple(env.obs.shape), dtype=torch.uint8, device=dev)
    actions = torch.zeros(n, dtype=torch.int32, device=dev)
    vp = lambda t: C.c_void_p(t.data_ptr())
    p_act, p_rew, p_done, p_life, p_legal = vp(actions), vp(env.reward), vp(env.done), vp(env.lifetime), vp(env.legal_mask)
    p_ring = [C.c_void_p(ring[s].data_ptr()) for s in range(RING)]

    def pair(slot, stream):
        # random-legal pick + env step in ONE launch (dq_env_step_random; tests/test_env_gpu.py checks it equals
        # dq_policy_random_legal_next followed by dq_env_step bit for bit)
        _lib.check(L.dq_env_step_random(h, p_ring[slot], p_rew, p_done, p_life, p_legal, p_act, 1, stream))

    cur = lambda: C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    env.reset()
    _lib.check(L.dq_policy_seek(h, 0, cur()))
    torch.cuda.synchronize()

    # ---- the timed path: multi-step rollout launches (dq_env_rollout_random, ROLL steps of every lattice per launch; bit-identical
    #      to ROLL single-step launches, tests/test_env_gpu.py).  Step s writes its observations into ring slot (cursor+s) % RING
    #      and row s of the per-step outputs, so every step still produces every output in HBM.
    roll_out = [torch.empty((ROLL, n), dtype=dt, device=dev) for dt in (torch.float32, torch.uint8, torch.int32, torch.int32)]
    roll_legal = torch.empty((ROLL, n, env.mask_words), dtype=torch.int64, device=dev)
    state = {"cursor": 0, "launches": 0}

    def rollout(steps):
        _lib.check(L.dq_env_rollout_random(h, steps, vp(ring), RING, state["cursor"], vp(roll_out[0]), vp(roll_out[1]), vp(roll_out[2]),
                                           vp(roll_legal), vp(roll_out[3]), 1, cur()))
        state["cursor"] = (state["cursor"] + steps) % RING
        state["launches"] += 1

    def run_steps(k):
        full, rest = divmod(k, ROLL)
        for _ in range(full):
            rollout(ROLL)
        if rest:
            rollout(rest)

    # the older launch shape, kept as a second number: a CUDA graph of RING single-step launches
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for s in range(RING):
            pair(s, cur())                      # warm the kernels before capture
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        for s in range(RING):
            pair(s, cur())

    def run_single_steps(k):
        full, rest = divmod(k, RING)
        for _ in range(full):
            graph.replay()
        for s in range(rest):
            pair(s, cur())

    sampler = ClockSampler(local) if rank == 0 else None
    run_steps(Wm)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    state["launches"] = 0
    e0.record()
    run_steps(K)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_wall1 = time.perf_counter()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    value = world * n * K / (ms * 1e-3)
    timed_launches = state["launches"]
    # second number: the same K steps as single-step launches (one launch per step, replayed from a CUDA graph)
    run_single_steps(RING)
    torch.cuda.synchronize()
    e0.record()
    run_single_steps(K)
    e1.record()
    torch.cuda.synchronize()
    ms_single = e0.elapsed_time(e1)
    single = {"value": n * K / (ms_single * 1e-3), "unit": UNIT + " on this rank", "ms_per_step": ms_single / K,
              "launch": "CUDA graph of %d single-step launches (dq_env_step_random)" % RING}

    # ---- roofline: per-launch duration of the env-step kernel, CUDA events around every rollout launch.
    # The launches are queued behind a ~25 ms device-side delay so that, when the GPU reaches them,
    # they run back to back and the event pairs bracket device time only (not Python launch gaps).
    roof = None
    if rank == 0:
        nprof = 24
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(nprof)]
        torch.cuda._sleep(int(25e-3 * 1.9e9))
        for i in range(nprof):
            evs[i][0].record()
            rollout(ROLL)
            evs[i][1].record()
        torch.cuda.synchronize()
        durs = sorted(a.elapsed_time(b) * 1e-3 for a, b in evs)
        mean_s = sum(durs) / len(durs)
        peak, peak_src = measured_peak()
        achieved = ROLL * n * BYTES_PER_STEP / mean_s / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "env_step_traffic.json")
        if os.path.exists(tpath) and args.workload == "c3":
            try:
                traffic = json.load(open(tpath)).get("dram_bytes_per_step") * ROLL      # ncu capture of one rollout launch, per step
            except Exception:
                traffic = None
        roof = {"bound": "hbm", "kernel": "env_step_kernel<%d,false>" % D, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "kernel_us_mean": mean_s * 1e6, "kernel_us_median": durs[len(durs) // 2] * 1e6,
                "algorithmic_bytes_per_launch": ROLL * n * BYTES_PER_STEP, "launches_timed": nprof, "steps_per_launch": ROLL,
                "kernel_us_per_step": mean_s * 1e6 / ROLL}

    # ---- the same kernel at larger lattice counts (extra evidence): at C3's 16 384 lattices a launch is bounded by the
    #      latency of one CTA's dependent chain; the sweep shows where the kernel goes once a launch has enough tiles
    scaling = None
    if rank == 0 and not args.no_dqn:
        scaling = []
        peak, _ = measured_peak()
        for nn in (16384, 65536, 262144, 1048576):
            e2 = VecSurfaceCodeEnv(D, P, P, MODEL, USE_Y, VD, None, n_envs=nn, seed=SEED + 5, env_id_base=0, device=dev)
            nbuf = max(2, int(300e6 // (nn * e2.obs[0].numel())) + 1)          # rotate observation slots past the L2 size
            e2.reset()
            ring2 = torch.zeros((nbuf,) + tuple(e2.obs.shape), dtype=torch.uint8, device=dev)
            S2 = 16
            o2 = [torch.empty((S2, nn), dtype=dt, device=dev) for dt in (torch.float32, torch.uint8, torch.int32, torch.int32)]
            l2 = torch.empty((S2, nn, e2.mask_words), dtype=torch.int64, device=dev)

            def roll2(i):
                _lib.check(L.dq_env_rollout_random(e2._h, S2, vp(ring2), nbuf, (i * S2) % nbuf, vp(o2[0]), vp(o2[1]), vp(o2[2]), vp(l2),
                                                   vp(o2[3]), 1, cur()))
            for i in range(2):
                roll2(i)
            torch.cuda.synchronize()
            a_ev, b_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 4
            a_ev.record()
            for i in range(reps):
                roll2(i)
            b_ev.record()
            torch.cuda.synchronize()
            t = a_ev.elapsed_time(b_ev) * 1e-3 / (reps * S2)
            scaling.append({"lattices": nn, "us_per_step": t * 1e6, "env_steps_per_s": nn / t, "steps_per_launch": S2,
                            "achieved_GBps": nn * BYTES_PER_STEP / t / 1e9, "frac": nn * BYTES_PER_STEP / t / 1e9 / peak})
            del ring2, o2, l2
            e2.close()

    # ---- e2e: host buffers through dq_env_step_host (H2D actions, D2H every output, every step)
    ke = min(K, 64)
    rng = np.random.default_rng(SEED + rank)
    host_actions = torch.from_numpy(rng.integers(0, env.num_actions, size=(ke + 3, n), dtype=np.int32)).pin_memory()
    hb = env._host_buffers()
    hp = lambda t: C.c_void_p(t.data_ptr())
    torch.cuda.synchronize()

    def host_step(i):
        _lib.check(L.dq_env_step_host(h, C.c_void_p(host_actions[i].data_ptr()), hp(hb["obs"]), hp(hb["reward"]),
                                      hp(hb["done"]), hp(hb["lifetime"]), hp(hb["legal"]), 1))
    for i in range(3):
        host_step(i)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(3, ke + 3):
        host_step(i)
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_value = world * n * ke / float(dt.item())
    h2d = n * 4
    d2h = n * (env.obs[0].numel() + 4 + 1 + 4 + 8 * env.mask_words)
    host_expand = False
    try:                         # DQ_HOST_EXPAND=1 (opt-in): same call, observations cross PCIe bit-packed and are expanded by host threads
        he = C.c_int64(0)
        if L.dq_env_info(h, 10, C.byref(he)) == 0 and he.value == 1:
            host_expand = True
            d2h = int((env.state_words - 7) * env.state_stride * 8 + n * (4 + 1 + 4 + 8 * env.mask_words))
    except Exception:            # noqa: BLE001
        pass

    # ---- the same loop with the observations returned PACKED (one bit per cell, the rows the Q-network consumes):
    #      an extra line of evidence next to e2e, never a replacement for it; a failure here must not cost the bench line
    e2e_packed, packed_err, packed_dt, packed_bytes = None, None, -1.0, 0
    if world > 1:
        dist.barrier()
    try:                         # no collective inside: a failure on one rank must not leave the others waiting
        pk = env._packed_host_buffer()
        packed_bytes = int(pk.numel() * 8 + n * (4 + 1 + 4 + 8 * env.mask_words))

        def host_step_packed(i):
            _lib.check(L.dq_env_step_host_packed(h, C.c_void_p(host_actions[i].data_ptr()), hp(pk), hp(hb["reward"]),
                                                 hp(hb["done"]), hp(hb["lifetime"]), hp(hb["legal"]), 1))
        for i in range(3):
            host_step_packed(i)
        t0 = time.perf_counter()
        for i in range(3, ke + 3):
            host_step_packed(i)
        packed_dt = time.perf_counter() - t0
    except Exception as ex:      # noqa: BLE001 -- reported, not fatal
        packed_err = "%s: %s" % (type(ex).__name__, ex)
    dtp = torch.tensor([packed_dt, -packed_dt], dtype=torch.float64, device=dev)       # max over ranks of (dt, -dt): slowest, and any failure (-1)
    if world > 1:
        dist.all_reduce(dtp, op=dist.ReduceOp.MAX)
    if packed_err is None and float(dtp[1].item()) < 0:
        e2e_packed = {"value": world * n * ke / float(dtp[0].item()), "unit": UNIT, "h2d_bytes_per_step": h2d,
                      "d2h_bytes_per_step": packed_bytes, "steps": ke,
                      "api": "dq_env_step_host_packed (observations as bit-packed rows uint64 [C*PW][stride]; "
                             "envs.unpack_observations expands them on the host when a caller needs bytes)"}
    else:
        e2e_packed = {"error": packed_err or "failed on another rank"}

    # ---- DQN inner loop on the same lattices (extra evidence, not the headline metric):
    #   act:   Q(s) for every lattice from the packed rows in the env state -> eps-greedy pick -> env step (no byte boards)
    #   train: act + store in the replay ring + one double-DQN update (batch 4096) per iteration
    dqn = None
    if rank == 0 and not args.no_dqn:
        from deepq_decoding_b200 import agents as A
        spec = A.build_convolutional_nn([[64, 3, 2], [32, 2, 1], [32, 2, 1]], [[512, 0.2]], (7, 11, 11), env.num_actions)
        agent = A.DQNAgent(model=spec, nb_actions=env.num_actions, memory=A.SequentialMemory(limit=64 * n), nb_steps_warmup=0,
                           target_model_update=10 ** 9, policy=A.EpsGreedyQPolicy(eps=0.1), test_policy=A.GreedyQPolicy(masked_greedy=True),
                           enable_dueling_network=True, batch_size=4096, seed=SEED, device=dev)
        agent.compile(A.Adam(lr=1e-5), max_envs=n)
        rows_ptr, nrows, stride = C.c_void_p(), C.c_int64(), C.c_int64()
        _lib.check(L.dq_env_packed_obs(h, C.byref(rows_ptr), C.byref(nrows), C.byref(stride)))
        from deepq_decoding_b200.qnet import device_view
        rows_view = device_view(rows_ptr.value, (nrows.value, stride.value), "<i8", dev)
        rring = A.ReplayRing(65, nrows.value, stride.value, n, dev)
        st = cur()

        def act_iter(i, store):
            if store:
                rring.push_obs(rows_view)
            a = agent._act(env, rows_ptr.value, i, 0.1, True)
            _lib.check(L.dq_env_step(h, vp(a), None, p_rew, p_done, p_life, p_legal, 1, cur()))
            if store:
                rring.push_outcome(a, env.reward, env.done)
            return a

        def timed(fn, iters):
            for i in range(5):
                fn(i)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(iters):
                fn(5 + i)
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) * 1e-3 / iters

        t_fwd = timed(lambda i: agent.model.forward_packed(rows_ptr.value, stride.value, n), 30)
        t_fwd_tc = timed(lambda i: agent.model.forward_packed(rows_ptr.value, stride.value, n, precision="bf16"), 30)
        t_act32 = timed(lambda i: act_iter(i, False), 40)
        agent.act_precision = "bf16"
        t_act = timed(lambda i: act_iter(i, False), 60)

        def train_iter(i):
            act_iter(i, True)
            if rring.filled >= 1:
                agent.train_on_ring(rring, i)
        t_train = timed(train_iter, 40)
        flops = agent.model.flops_per_sample
        tf_peak = 1393.5
        try:
            tf_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"])
        except Exception:
            pass
        dqn = {"act_env_steps_per_s": n / t_act, "act_ms_per_iteration": t_act * 1e3,
               "train_env_steps_per_s": n / t_train, "train_ms_per_iteration": t_train * 1e3,
               "train_batch": 4096, "updates_per_iteration": 1,
               "act_fp32_env_steps_per_s": n / t_act32,
               "qnet_forward_fp32_ms": t_fwd * 1e3, "qnet_forward_fp32_tflops": flops * n / t_fwd / 1e12,
               "qnet_forward_bf16_ms": t_fwd_tc * 1e3, "qnet_forward_bf16_tflops": flops * n / t_fwd_tc / 1e12,
               "qnet_frac_of_bf16_sustained_peak": flops * n / t_fwd_tc / 1e12 / tf_peak,
               "cpu_serial_loop": cpu_serial_dqn_loop(min(6.0, max(0.5, args.cpu_seconds / 2))),
               "qnet_precision": "acting: bf16 tcgen05 (fp32 accumulate in TMEM); updates: fp32 SIMT",
               "qnet_flops_per_sample": flops, "policy": "eps-greedy 0.1 over legal actions, masked greedy"}

    # ---- logical error rate (second half of the BASELINE metric): the reference's published d5_dp/0.007 agent
    #      (weights carried as a test fixture) evaluated greedily here, one episode per lattice, LER := 1 / <lifetime>
    ler = None
    wpath = os.path.join(ROOT, "tests", "golden", "dqn_d5_dp_0.007.npz")
    if rank == 0 and not args.no_dqn and os.path.exists(wpath):
        from deepq_decoding_b200 import agents as A
        z = np.load(wpath)
        ev_env = VecSurfaceCodeEnv(D, P, P, MODEL, USE_Y, VD, None, n_envs=n, seed=SEED + 1, env_id_base=10 * n, device=dev)
        ev = A.DQNAgent(model=A.build_convolutional_nn([[64, 3, 2], [32, 2, 1], [32, 2, 1]], [[512, 0.2]], (7, 11, 11), ev_env.num_actions),
                        nb_actions=ev_env.num_actions, memory=A.SequentialMemory(limit=100), test_policy=A.GreedyQPolicy(masked_greedy=True),
                        enable_dueling_network=True, device=dev, act_precision="bf16")
        ev.compile(A.Adam(lr=1e-5), max_envs=n)
        ev.model.set_keras_weights([(z["conv%d_k" % i], z["conv%d_b" % i]) for i in range(3)], [(z["dense%d_k" % i], z["dense%d_b" % i]) for i in range(3)])
        t0 = time.perf_counter()
        life = np.array(ev.test(ev_env, nb_episodes=n, verbose=0).history["episode_lifetime"], dtype=np.float64)
        ler = {"agent": "reference trained_models/d5_dp/0.007/final_dqn_weights.h5f (tests/golden fixture), greedy masked policy, bf16 acting",
               "episodes": int(len(life)), "mean_lifetime_cycles": float(life.mean()), "standard_error": float(life.std() / np.sqrt(len(life))),
               "logical_error_rate_per_cycle": float(1.0 / life.mean()), "reference_published_mean_lifetime": 270.42,
               "eval_seconds": time.perf_counter() - t0,
               "trained_here": "profiles/r1_train_curriculum_dp_p007.json: 342.9 +- 3.7 cycles after 109 s of training from scratch"}
        ev_env.close()

    experiments = None
    if rank == 0 and world == 1 and args.workload == "c3" and not args.no_experiments:
        try:
            torch.cuda.synchronize()
            experiments = run_experiments()
        except Exception as ex:      # noqa: BLE001
            experiments = {"error": str(ex)[:200]}
    if rank == 0:
        cb, _, _ = cpu_oracle_run(args.cpu_seconds)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u64", "data": "synthetic",
                "config": {"workload": WORKLOAD, "lattices_total": world * n,
                           "l2": "each step writes its %.1f MB of observations into one of %d ring slots (%.0f MB > L2), "
                                 "so no step's writes are absorbed by the previous step's lines" % (
                                     ring[0].numel() / 1e6, RING, ring.numel() / 1e6),
                           "launch": "rollout launches of %d steps each (dq_env_rollout_random: random-legal pick + env step, every lattice "
                                     "advanced %d steps per launch; bit-identical to single-step launches)" % (ROLL, ROLL),
                           "parallelism": "lattices sharded by rank, no data-path collective"},
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "steps": ke, "api": "dq_env_step_host (pinned host actions in, all outputs to pinned host buffers)",
                        "host_expand": host_expand,
                        "policy": "uniform random action indices pre-generated on the host"},
                "e2e_packed": e2e_packed,
                "gpu_launches": timed_launches, "single_step_launches": single,
                "roofline": roof, "roofline_scaling": scaling, "cpu_baseline": cb, "dqn": dqn, "logical_error_rate": ler,
                "experiments": experiments}
        print(json.dumps(line), flush=True)
    env.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------- opt-in builds, timed beside the default
def experiment_child(args):
    """One opt-in build / mode in its own process (the library and the DQ_* switches are chosen per process): a fixed seeded
    rollout whose outputs are folded into a checksum (must equal the default build's), then the rollout and host-buffer rates."""
    import numpy as np
    import torch
    from deepq_decoding_b200 import _lib
    from deepq_decoding_b200.envs import VecSurfaceCodeEnv
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    L = _lib.lib()
    n = N_PER_GPU
    vp = lambda t: C.c_void_p(t.data_ptr())
    cur = lambda: C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    env = VecSurfaceCodeEnv(D, P, P, MODEL, USE_Y, VD, None, n_envs=n, seed=SEED + 77, env_id_base=0, device=dev)
    ring = torch.zeros((RING,) + tuple(env.obs.shape), dtype=torch.uint8, device=dev)
    S = 64
    out = [torch.empty((S, n), dtype=dt, device=dev) for dt in (torch.float32, torch.uint8, torch.int32, torch.int32)]
    legal = torch.empty((S, n, env.mask_words), dtype=torch.int64, device=dev)
    env.reset()
    _lib.check(L.dq_policy_seek(env._h, 0, cur()))

    def roll(i):
        _lib.check(L.dq_env_rollout_random(env._h, S, vp(ring), RING, (i * S) % RING, vp(out[0]), vp(out[1]), vp(out[2]), vp(legal), vp(out[3]), 1, cur()))
    roll(0)
    torch.cuda.synchronize()
    mix = lambda t: int((t.to(torch.int64).flatten() * (torch.arange(t.numel(), device=dev, dtype=torch.int64) % 1000003 + 1)).sum().item())
    checksum = [mix(out[0]), mix(out[1]), mix(out[2]), mix(out[3]), mix(legal), mix(ring.view(torch.uint8)), mix(env.get_state_words())]
    for i in range(1, 4):
        roll(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 24
    a.record()
    for i in range(reps):
        roll(4 + i)
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1e3 / (reps * S)
    res = {"checksum": checksum, "rollout_us_per_step": us, "env_steps_per_s": n / us * 1e6, "steps_per_launch": S}
    # host-buffer rate (dq_env_step_host), as the e2e leg does it
    rng = np.random.default_rng(SEED)
    ke = 24
    host_actions = torch.from_numpy(rng.integers(0, env.num_actions, size=(ke + 3, n), dtype=np.int32)).pin_memory()
    hb = env._host_buffers()
    hp = lambda t: C.c_void_p(t.data_ptr())
    def host_step(i):
        _lib.check(L.dq_env_step_host(env._h, C.c_void_p(host_actions[i].data_ptr()), hp(hb["obs"]), hp(hb["reward"]),
                                      hp(hb["done"]), hp(hb["lifetime"]), hp(hb["legal"]), 1))
    for i in range(3):
        host_step(i)
    t0 = time.perf_counter()
    for i in range(3, ke + 3):
        host_step(i)
    dt = time.perf_counter() - t0
    res["host_env_steps_per_s"] = n * ke / dt
    res["host_obs_checksum"] = mix(hb["obs"].to(dev))
    he = C.c_int64(0)
    res["host_expand"] = bool(L.dq_env_info(env._h, 10, C.byref(he)) == 0 and he.value == 1)
    env.close()
    print("EXPERIMENT " + json.dumps(res), flush=True)


def run_experiments():
    """Times the opt-in builds that were written without GPU access next to the default build, each in a child process with a
    time limit; nothing here feeds `value` / `e2e`.  A child that fails or is missing its library is reported, not fatal."""
    arms = [("default", {}),
            ("stream_obs", {"DQ_DECODING_LIB": os.path.join(ROOT, "build", "variants", "libdq_so.so")}),
            ("defer1_stream_obs", {"DQ_DECODING_LIB": os.path.join(ROOT, "build", "variants", "libdq_dfso.so")}),
            ("host_expand", {"DQ_HOST_EXPAND": "1"}),
            ("defer2_stream_obs", {"DQ_DECODING_LIB": os.path.join(ROOT, "build", "variants", "libdq_df2so.so")})]    # last: the one with a new barrier
    res, t_start = {}, time.perf_counter()
    for name, extra in arms:
        if time.perf_counter() - t_start > 150:          # the whole leg stays within a few minutes whatever happens
            res[name] = {"skipped": "time budget of the experiments leg spent"}
            continue
        libp = extra.get("DQ_DECODING_LIB")
        if libp and not os.path.exists(libp):
            res[name] = {"skipped": "library not built"}
            continue
        try:
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--experiment-child"], env=dict(os.environ, **extra),
                                 capture_output=True, text=True, timeout=60)
            line = [l for l in out.stdout.splitlines() if l.startswith("EXPERIMENT ")]
            res[name] = json.loads(line[-1][len("EXPERIMENT "):]) if line else {"error": (out.stderr or out.stdout)[-300:]}
        except Exception as ex:      # noqa: BLE001 -- reported, not fatal
            res[name] = {"error": "%s: %s" % (type(ex).__name__, str(ex)[:200])}
    ref = res.get("default", {})
    for name, r in res.items():
        if "checksum" in r and "checksum" in ref:
            r["identical_to_default_build"] = r["checksum"] == ref["checksum"] and r["host_obs_checksum"] == ref["host_obs_checksum"]
    for r in res.values():
        r.pop("checksum", None)
    return res


def select_workload(name):
    global D, VD, P, MODEL, N_PER_GPU, BYTES_PER_STEP, WORKLOAD, METRIC
    D, VD, P, MODEL, N_PER_GPU, BYTES_PER_STEP = WORKLOADS[name]
    if name != "c3":
        WORKLOAD = "%s: d=%d %s p_phys=p_meas=%g volume_depth=%d use_Y=False, %d lattices/GPU, random-legal policy, referee = %s" % (
            name.upper(), D, MODEL, P, VD, N_PER_GPU, "shipped nn_d5_X_p5 tabulated" if D == 5 else "minimum-weight table")
        METRIC = "env-steps/sec at d=%d %s p=%g" % (D, "depolarising" if MODEL == "DP" else "bit-flip", P)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=32768)
    ap.add_argument("--warmup", type=int, default=64)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="wall budget of the cpu_baseline leg")
    ap.add_argument("--no-dqn", action="store_true", help="skip the DQN inner-loop measurements")
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS), help="BASELINE.json config (default c3 = the metric's)")
    ap.add_argument("--experiment-child", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--no-experiments", action="store_true", help="skip the child processes that time the opt-in builds")
    args = ap.parse_args()
    select_workload(args.workload)
    if args.experiment_child:
        return experiment_child(args)
    if args.workload != "c3":
        args.no_dqn = True
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
