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
OBS_BYTES = 847               # C * (2d+1)^2 of the selected workload
BYTES_PER_STEP = 996          # obs 847 + action 4 + reward 4 + done 1 + lifetime 4 + mask 8 + 2 x 64 state
WORKLOAD = ("C3: d=5 DP p_phys=p_meas=0.007 volume_depth=5 use_Y=False, %d lattices/GPU, random-legal policy, "
            "referee = shipped nn_d5_DP_p5 tabulated" % N_PER_GPU)


def bench_config(world):
    """The `config` object of the JSON line: the same dict from both arms (the reference arm runs on this arm's config)."""
    n = N_PER_GPU
    return {"workload": WORKLOAD, "lattices_total": world * n, "lattices_per_gpu": n,
            "l2": "GPU arm: each step writes its %.1f MB of observations into one of %d ring slots (%.0f MB > L2), so no step's writes are "
                  "absorbed by the previous step's lines" % (n * OBS_BYTES / 1e6, RING, RING * n * OBS_BYTES / 1e6),
            "launch": "GPU arm: the K timed steps run as dq_env_rollout_random launches of up to %d steps of every lattice (random-legal pick + "
                      "env step per step; bit-identical to single-step launches), queued behind a 2 ms device-side delay so the event pair "
                      "brackets device time; the launches actually made are listed in launch_detail" % ROLL,
            "parallelism": "lattices sharded by rank, no data-path collective (the gradient exchange of training is measured in dqn_dp)"}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------- CPU arm
def cpu_oracle_run(budget_s, n=None):
    """Times the oracle on all host cores: random-legal policy + step, like the GPU loop (n = lattices of the selected workload)."""
    import numpy as np
    n = N_PER_GPU if n is None else n
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
    cb = dict(value=n * steps / el, unit=UNIT, cores=cores, kind="port",
              sample="%d vectorised steps of %d lattices (%.1f s wall) of the same workload, oracle/dq_oracle.c "
                     "with OpenMP over lattices" % (steps, n, el))
    rp = reference_python_result()
    if rp is not None:
        cb["reference_python"] = rp
    return cb, n * steps / el, el / steps


def reference_python_result():
    """The UNMODIFIED Python reference timed on host cores (tools/time_reference_cpu.py).  It needs /root/reference, which exists
    only in the build container, so the committed result of that run is reported next to the live C-port number."""
    path = os.path.join(ROOT, "profiles", "reference_cpu_build_container.json")
    try:
        r = json.load(open(path))
        return {"value": r["env_steps_per_s_aggregate"], "unit": UNIT, "cores": r["processes"], "per_core": r["env_steps_per_s_per_process_mean"],
                "seconds_per_process": r["seconds_per_process"], "host_cpu": r["host"]["cpu"],
                "source": "profiles/reference_cpu_build_container.json (tools/time_reference_cpu.py, build container; not timed in this run)",
                "what": r["what"]}
    except Exception:
        return None


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
            "config": bench_config(args.gpus),
            "reference_note": "CPU oracle port of the reference env on the host cores (all of them, whatever --gpus says); each step is "
                              "one vectorised step of %d lattices; the run is time-bounded, not step-bounded" % N_PER_GPU,
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
def e2e_legs(L, _lib, torch, np, dist, dev, world, rank, n, K):
    """Host-buffer throughput: every step copies the actions in from pinned host memory and every output back to pinned host
    memory.  The policy is the same random-legal pick as on the device path, computed on the host from the legal masks that came
    back (dq_policy_random_legal_host), so the workload mix equals the resident loop's and the CPU arm's."""
    from deepq_decoding_b200.envs import VecSurfaceCodeEnv
    # Host-side loops are exposed to whatever else the box's cores are doing: each leg times BLOCKS consecutive blocks of `ke` steps
    # (barrier before each, max over ranks per block) and reports the MEDIAN block; every block's rate is listed.
    ke = max(20, min(K, 64))
    BLOCKS = 5
    half = n // 2
    envs = [VecSurfaceCodeEnv(D, P, P, MODEL, USE_Y, VD, None, n_envs=cnt, seed=SEED + 3, env_id_base=rank * n + base, device=dev)
            for base, cnt in ((0, half), (half, n - half))]
    whole = VecSurfaceCodeEnv(D, P, P, MODEL, USE_Y, VD, None, n_envs=n, seed=SEED + 3, env_id_base=rank * n, device=dev)
    he = C.c_int64(0)
    host_expand = bool(L.dq_env_info(whole._h, 10, C.byref(he)) == 0 and he.value == 1)
    small = 4 + 1 + 4 + 8 * whole.mask_words
    packed_bytes = int((whole.state_words - 7) * whole.state_stride * 8)
    obs_bytes = int(whole.obs[0].numel())

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def timed_blocks(step, first):
        """step(i) = one full step of every lattice of this rank, results landed in host memory"""
        rates, i = [], first
        for _ in range(BLOCKS):
            sync_all()
            t0 = time.perf_counter()
            for _ in range(ke):
                step(i)
                i += 1
            t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            rates.append(world * n * ke / float(t.item()))
        return sorted(rates)[BLOCKS // 2], rates

    timing = "median of %d consecutive blocks of %d steps (wall clock around each block, max over ranks); block_values lists them all" % (BLOCKS, ke)
    # (a) the public call on one handle of n lattices: random-legal pick on the host, dq_env_step_host, every output landed -- per step
    whole.reset_host()

    def one_handle_step(i):
        whole.random_legal_actions_host(i); whole.step_host_begin(); whole.step_host_end()
    for i in range(6):
        one_handle_step(i)
    val, rates = timed_blocks(one_handle_step, 6)
    e2e = {"value": val, "unit": UNIT, "h2d_bytes_per_step": n * 4,
           "d2h_bytes_per_step": n * small + (packed_bytes if host_expand else n * obs_bytes), "steps": ke * BLOCKS, "timing": timing,
           "block_values": rates, "host_bytes_delivered_per_step": n * (obs_bytes + small),
           "api": "VecSurfaceCodeEnv.random_legal_actions_host + step_host_begin / step_host_end (dq_policy_random_legal_host, dq_env_step_host) on one "
                  "handle of %d lattices: uint8 observations [N,C,H,H], reward, done, lifetime and legal masks land in pinned host memory every step" % n,
           "host_expand": host_expand, "host_threads": os.environ.get("DQ_HOST_THREADS", "default (usable CPUs - 1)"),
           "how": ("the bitmap rows cross PCIe bit-packed (DMA) and the library's host threads expand them into the byte observations; "
                   "the kernel reads the actions from, and stores reward / done / lifetime / legal masks into, the caller's pinned buffers "
                   "(zero-copy: those bytes cross PCIe inside the timed region by the kernel's own loads and stores; DQ_HOST_ZEROCOPY=0 for DMA copies)")
                  if host_expand else "the kernel writes bytes; all of them are copied back",
           "policy": "uniform random-legal, computed on the host from the returned legal masks (same picks as the device policy)"}
    # (b) two handles of n/2 lattices driven alternately -- begin(A) begin(B) end(A) policy(A) begin(A) end(B) ... -- so that one handle's
    #     kernel and copies run under the other's host-side work
    try:
        for e in envs:
            e.reset_host()
        step = [0, 0]

        def begin(k):
            envs[k].random_legal_actions_host(step[k])
            envs[k].step_host_begin()
            step[k] += 1
        for k in (0, 1):
            begin(k)

        def two_handle_step(i):
            for k in (0, 1):
                envs[k].step_host_end(); begin(k)
        for i in range(6):
            two_handle_step(i)
        val2, rates2 = timed_blocks(two_handle_step, 0)
        for k in (0, 1):
            envs[k].step_host_end()
        e2e["two_handles"] = {"value": val2, "unit": UNIT, "block_values": rates2,
                              "api": "the same calls on two handles of %d lattices, split begin / end, driven alternately" % half}
    except Exception as ex:      # noqa: BLE001
        e2e["two_handles"] = {"error": "%s: %s" % (type(ex).__name__, ex)}
    # (c) observations returned PACKED (one bit per cell, the rows the Q-network consumes): nothing to expand anywhere
    e2e_packed = None
    try:
        hb, pk = whole._host_buffers(), whole._packed_host_buffer()
        hp = lambda t: C.c_void_p(t.data_ptr())

        def packed_step(i):
            whole.random_legal_actions_host(1000 + i)
            _lib.check(L.dq_env_step_host_packed(whole._h, hp(hb["actions"]), hp(pk), hp(hb["reward"]), hp(hb["done"]), hp(hb["lifetime"]), hp(hb["legal"]), 1))
        for i in range(4):
            packed_step(i)
        val3, rates3 = timed_blocks(packed_step, 4)
        e2e_packed = {"value": val3, "unit": UNIT, "h2d_bytes_per_step": n * 4,
                      "d2h_bytes_per_step": packed_bytes + n * small, "steps": ke * BLOCKS, "block_values": rates3,
                      "api": "dq_env_step_host_packed (observations as bit-packed rows uint64 [C*PW][stride]; envs.unpack_observations expands "
                             "them when a caller needs bytes), same policy"}
    except Exception as ex:      # noqa: BLE001 -- reported, not fatal (no collective between here and the next barrier on the success path either)
        e2e_packed = {"error": "%s: %s" % (type(ex).__name__, ex)}
    for e in envs + [whole]:
        e.close()
    return e2e, e2e_packed


def dqn_data_parallel_leg(L, _lib, torch, np, dist, dev, world, rank, n):
    """The one collective of the path: data-parallel DQN updates over the sharded lattices, on EVERY rank.
      * per-update device time of the gradient exchange + Adam: the fused peer-memory kernel (csrc/dq_comm.cu) vs NCCL all-reduce + Adam kernel;
      * a short sharded `DQNAgent.fit` (C4 shape: n lattices per GPU) through collective='fused' and through 'nccl': parameters must stay
        bit-identical across the ranks, and the two collectives must agree at world 2 (one fp32 add either way);
      * sharded greedy evaluation of the reference's published d5_dp/0.007 agent, merged with parallel.reduce_lifetimes: LER at N GPUs."""
    from deepq_decoding_b200 import agents as A, parallel
    from deepq_decoding_b200.envs import VecSurfaceCodeEnv
    out = {"world": world, "lattices_per_gpu": n}
    st = lambda: C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())
    group = dist.group.WORLD if world > 1 else None
    nparams = 193283 if (D, MODEL) == (5, "DP") else None
    # ---- (1) exchange + Adam per update, device time, max over ranks
    opt = A.Adam(lr=1e-5)

    def timed(fn, iters=200, warm=20):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) * 1e3
    npar = nparams or 193283
    params, m, v, g = torch.randn(npar, device=dev), torch.zeros(npar, device=dev), torch.zeros(npar, device=dev), torch.randn(npar, device=dev)
    tcount = [0]
    if world > 1:
        comm = parallel.FusedAllreduceAdam(npar, dev, group)

        def fused():
            tcount[0] += 1
            comm.grads()
            comm.step(params, m, v, opt, tcount[0], st())

        def nccl():
            tcount[0] += 1
            dist.all_reduce(g)
            _lib.check(L.dq_adam_step(p(params), p(m), p(v), p(g), npar, opt.lr, opt.beta_1, opt.beta_2, opt.epsilon, tcount[0], 1.0 / world, st()))
        out["fused_us"], out["nccl_us"] = timed(fused), timed(nccl)
        comm.check(); comm.close()
    else:
        # one GPU: the same kernel with itself as only peer (world 1) against the stand-alone Adam kernel
        h = C.c_void_p()
        _lib.check(L.dq_comm_create(C.byref(h), 0, 1, npar, dev.index or 0))

        def fused():
            tcount[0] += 1
            gp = C.c_void_p()
            _lib.check(L.dq_comm_next_grads(h, C.byref(gp)))
            _lib.check(L.dq_comm_allreduce_adam(h, p(params), p(m), p(v), opt.lr, opt.beta_1, opt.beta_2, opt.epsilon, tcount[0], st()))

        def adam_only():
            tcount[0] += 1
            _lib.check(L.dq_adam_step(p(params), p(m), p(v), p(g), npar, opt.lr, opt.beta_1, opt.beta_2, opt.epsilon, tcount[0], 1.0, st()))
        out["fused_us"], out["nccl_us"] = timed(fused), None
        out["adam_kernel_only_us"] = timed(adam_only)
        flag = C.c_int(0)
        _lib.check(L.dq_comm_status(h, C.byref(flag)))
        out["note"] = "one GPU: dq_comm_allreduce_adam with world = 1 (its own region as only peer); there is nothing for NCCL to do"
        L.dq_comm_destroy(h)
    # ---- (2) short sharded fit through both collectives
    spec_args = ([[64, 3, 2], [32, 2, 1], [32, 2, 1]], [[512, 0.2]])

    def short_fit(collective, train_precision="fp32"):
        env = VecSurfaceCodeEnv(D, P, P, MODEL, USE_Y, VD, None, n_envs=n, seed=SEED + 21, env_id_base=rank * n, device=dev)
        spec = A.build_convolutional_nn(spec_args[0], spec_args[1], env.observation_space.shape, env.num_actions)
        pol = A.LinearAnnealedPolicy(A.EpsGreedyQPolicy(masked_greedy=False), attr="eps", value_max=1.0, value_min=0.05, value_test=0.0, nb_steps=20 * n)
        dqn = A.DQNAgent(model=spec, nb_actions=env.num_actions, memory=A.SequentialMemory(limit=16 * n), nb_steps_warmup=4 * n,
                         target_model_update=10 * n, policy=pol, test_policy=A.GreedyQPolicy(masked_greedy=True), gamma=0.99,
                         enable_dueling_network=True, batch_size=1024, seed=0, device=dev, process_group=group, collective=collective,
                         act_precision="bf16", target_precision=train_precision, train_precision=train_precision)
        dqn.compile(A.Adam(lr=1e-4), max_envs=n)
        if world > 1:
            parallel.broadcast_params_(dqn.model.params)
            dqn.target_params.copy_(dqn.model.params)
        t0 = time.perf_counter()
        dqn.fit(env, nb_steps=24 * n, verbose=0, episode_averaging_length=1000, success_threshold=1e9, stopping_patience=1e12)
        torch.cuda.synchronize()
        secs = time.perf_counter() - t0
        params = dqn.model.params.clone()
        env.close()
        return params, dqn.updates, secs
    pf, upd, secs = short_fit("fused")
    pn, _, _ = short_fit("nccl")
    ident = True
    if world > 1:
        lo, hi = pf.clone(), pf.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        ident = bool(torch.equal(lo.view(torch.int32), hi.view(torch.int32)))
        lo, hi = pn.clone(), pn.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        ident = ident and bool(torch.equal(lo.view(torch.int32), hi.view(torch.int32)))
    # the same fit with the updates on the tensor cores (bf16 forward / backward writing its gradient straight into the exchange region)
    try:
        pb, _, secs_bf16 = short_fit("fused", "bf16")
        ident_bf16 = True
        if world > 1:
            lo, hi = pb.clone(), pb.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            ident_bf16 = bool(torch.equal(lo.view(torch.int32), hi.view(torch.int32)))
        out.update({"fit_seconds_fused_bf16_updates": secs_bf16, "identical_bf16_updates": ident_bf16, "fit_finite_bf16_updates": bool(torch.isfinite(pb).all())})
    except Exception as ex:      # noqa: BLE001 -- reported in the line
        out["bf16_updates_error"] = "%s: %s" % (type(ex).__name__, str(ex)[:200])
    out.update({"identical": ident, "fit_updates": upd, "fit_env_steps": 24 * n * world, "fit_seconds_fused": secs,
                "fit_finite": bool(torch.isfinite(pf).all()),
                "fused_vs_nccl_max_abs_diff": float((pf - pn).abs().max()),
                "fit": "DQNAgent.fit on %d lattices per rank, batch 1024 per rank, one update per iteration after a warm-up of 4 iterations, 24 iterations" % n})
    # ---- (3) sharded greedy evaluation of the published agent
    wpath = os.path.join(ROOT, "tests", "golden", "dqn_d5_dp_0.007.npz")
    if os.path.exists(wpath) and (D, MODEL) == (5, "DP"):
        z = np.load(wpath)
        ev_env = VecSurfaceCodeEnv(D, P, P, MODEL, USE_Y, VD, None, n_envs=n, seed=SEED + 1, env_id_base=(10 + rank) * n, device=dev)
        ev = A.DQNAgent(model=A.build_convolutional_nn(spec_args[0], spec_args[1], (7, 11, 11), ev_env.num_actions),
                        nb_actions=ev_env.num_actions, memory=A.SequentialMemory(limit=100), test_policy=A.GreedyQPolicy(masked_greedy=True),
                        enable_dueling_network=True, device=dev, act_precision="bf16")
        ev.compile(A.Adam(lr=1e-5), max_envs=n)
        ev.model.set_keras_weights([(z["conv%d_k" % i], z["conv%d_b" % i]) for i in range(3)], [(z["dense%d_k" % i], z["dense%d_b" % i]) for i in range(3)])
        t0 = time.perf_counter()
        life = np.array(ev.test(ev_env, nb_episodes=n, verbose=0).history["episode_lifetime"], dtype=np.float64)
        mean, se, episodes = parallel.reduce_lifetimes(life, group)
        out["logical_error_rate"] = {"agent": "reference trained_models/d5_dp/0.007/final_dqn_weights.h5f (tests/golden fixture), greedy masked policy, bf16 acting",
                                     "episodes": episodes, "mean_lifetime_cycles": mean, "standard_error": se,
                                     "logical_error_rate_per_cycle": 1.0 / mean, "reference_published_mean_lifetime": 270.42,
                                     "within_3_se_of_published": bool(abs(mean - 270.42) <= 3 * se),
                                     # the published figure is itself a sample mean: all_results.p holds 270.42121212... = 446195 / 1650, i.e. at
                                     # most 1650 test episodes; with this run's spread its standard error is std / sqrt(1650)
                                     "published_episodes_at_most": 1650,
                                     "published_standard_error_at_least": float(se * np.sqrt(episodes / 1650.0)),
                                     "consistent_with_published_3_sigma": bool(abs(mean - 270.42) <= 3 * np.sqrt(se * se * (1.0 + episodes / 1650.0))),
                                     "eval_seconds": time.perf_counter() - t0, "merged_with": "parallel.reduce_lifetimes over %d rank(s)" % world}
        ev_env.close()
    return out


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
    # The CPU baseline runs BEFORE the ranks meet: the other ranks wait for rank 0 in the rendezvous (blocked on a socket), not in an
    # NCCL barrier that spins on the host cores the baseline is being timed on.
    cb = cpu_oracle_run(args.cpu_seconds)[0] if rank == 0 else None
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

    # ---- the timed path: multi-step rollout launches (dq_env_rollout_random, up to ROLL steps of every lattice per launch; bit-identical
    #      to single-step launches, tests/test_env_gpu.py).  Step s writes its observations into ring slot (cursor+s) % RING
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

    # the other launch shape, kept as a second number: a CUDA graph of RING single-step launches
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

    delay = lambda ms_: torch.cuda._sleep(int(ms_ * 1e-3 * 1.9e9))     # device-side spin: what is queued behind it starts back to back
    sampler = ClockSampler(local) if rank == 0 else None
    run_steps(Wm)
    run_steps(min(K, ROLL))                      # the launch shape of the timed region once more (its first launch of a given length is not a cold one)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    state["launches"] = 0
    delay(2.0)                                   # the K steps are queued while the device still spins: e0 -> e1 is device time, not host launch latency
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
    delay(2.0)
    e0.record()
    run_single_steps(K)
    e1.record()
    torch.cuda.synchronize()
    ms_single = e0.elapsed_time(e1)
    single = {"value": n * K / (ms_single * 1e-3), "unit": UNIT + " on this rank", "ms_per_step": ms_single / K,
              "launch": "CUDA graph of %d single-step launches (dq_env_step_random)" % RING}

    # ---- roofline: per-launch duration of the env-step kernel, CUDA events around every rollout launch of ROLL steps.
    # The launches are queued behind a ~25 ms device-side delay so that, when the GPU reaches them,
    # they run back to back and the event pairs bracket device time only (not Python launch gaps).
    roof = None
    if rank == 0:
        nprof = 24
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(nprof)]
        delay(25.0)
        for i in range(nprof):
            evs[i][0].record()
            rollout(ROLL)
            evs[i][1].record()
        torch.cuda.synchronize()
        durs = sorted(a.elapsed_time(b) * 1e-3 for a, b in evs)
        mean_s = sum(durs) / len(durs)
        peak, peak_src = measured_peak()
        achieved = ROLL * n * BYTES_PER_STEP / mean_s / 1e9
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "env_step_traffic.json")
        if os.path.exists(tpath) and args.workload == "c3":
            try:
                tj = json.load(open(tpath))
                traffic = tj.get("dram_bytes_per_step") * ROLL
                traffic_src = "stored ncu capture (%s): dram__bytes_read.sum + dram__bytes_write.sum of one rollout launch, per step, x %d; not measured in this run" % (
                    tj.get("capture", "profiles/env_step_traffic.json"), ROLL)
            except Exception:
                traffic = None
        roof = {"bound": "hbm", "kernel": "env_step_kernel<%d,false>" % D, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "kernel_us_mean": mean_s * 1e6, "kernel_us_median": durs[len(durs) // 2] * 1e6,
                "algorithmic_bytes_per_launch": ROLL * n * BYTES_PER_STEP, "launches_timed": nprof, "steps_per_launch": ROLL,
                "kernel_us_per_step": mean_s * 1e6 / ROLL, "value_over_kernel_rate": (value / world) / (n / (mean_s / ROLL))}

    def rollout_rate(e2, nn, S2, reps):
        nbuf = max(2, int(300e6 // (nn * e2.obs[0].numel())) + 1)          # rotate observation slots past the L2 size
        e2.reset()
        ring2 = torch.zeros((nbuf,) + tuple(e2.obs.shape), dtype=torch.uint8, device=dev)
        o2 = [torch.empty((S2, nn), dtype=dt, device=dev) for dt in (torch.float32, torch.uint8, torch.int32, torch.int32)]
        l2 = torch.empty((S2, nn, e2.mask_words), dtype=torch.int64, device=dev)

        def roll2(i):
            _lib.check(L.dq_env_rollout_random(e2._h, S2, vp(ring2), nbuf, (i * S2) % nbuf, vp(o2[0]), vp(o2[1]), vp(o2[2]), vp(l2),
                                               vp(o2[3]), 1, cur()))
        for i in range(2):
            roll2(i)
        torch.cuda.synchronize()
        a_ev, b_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        delay(2.0)
        a_ev.record()
        for i in range(reps):
            roll2(i)
        b_ev.record()
        torch.cuda.synchronize()
        return a_ev.elapsed_time(b_ev) * 1e-3 / (reps * S2)

    # ---- the same kernel at larger lattice counts, and on the other BASELINE configs (extra evidence on rank 0)
    scaling, others = None, None
    if rank == 0 and not args.no_dqn:
        scaling = []
        peak, _ = measured_peak()
        for nn in (16384, 65536, 262144, 1048576):
            e2 = VecSurfaceCodeEnv(D, P, P, MODEL, USE_Y, VD, None, n_envs=nn, seed=SEED + 5, env_id_base=0, device=dev)
            t = rollout_rate(e2, nn, 16, 4)
            scaling.append({"lattices": nn, "us_per_step": t * 1e6, "env_steps_per_s": nn / t, "steps_per_launch": 16,
                            "achieved_GBps": nn * BYTES_PER_STEP / t / 1e9, "frac": nn * BYTES_PER_STEP / t / 1e9 / peak})
            e2.close()
        if args.workload == "c3":
            others = {}
            for name in ("c5", "c2"):
                d_, vd_, p_, model_, nn, bytes_ = WORKLOADS[name]
                e2 = VecSurfaceCodeEnv(d_, p_, p_, model_, False, vd_, None, n_envs=nn, seed=SEED + 6, env_id_base=0, device=dev)
                t = rollout_rate(e2, nn, 256, 4)
                others[name] = {"workload": "d=%d %s p=%g volume_depth=%d, %d lattices per GPU" % (d_, model_, p_, vd_, nn),
                                "us_per_step": t * 1e6, "env_steps_per_s": nn / t, "steps_per_launch": 256,
                                "algorithmic_bytes_per_lattice_step": bytes_, "achieved_GBps": nn * bytes_ / t / 1e9, "frac": nn * bytes_ / t / 1e9 / peak}
                e2.close()

    # ---- e2e: host buffers in and out, every step
    e2e, e2e_packed = e2e_legs(L, _lib, torch, np, dist, dev, world, rank, n, K)

    # ---- the data-parallel DQN update on every rank (the path's one collective) + sharded evaluation of the published agent
    dqn_dp = None
    if not args.no_dqn and args.workload == "c3":
        try:
            dqn_dp = dqn_data_parallel_leg(L, _lib, torch, np, dist, dev, world, rank, n)
        except Exception as ex:      # noqa: BLE001 -- reported: the leg is the same code on every rank, so a failure is one on all of them
            dqn_dp = {"error": "%s: %s" % (type(ex).__name__, str(ex)[:300])}

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
        def ext_step(i):      # the launch every real policy uses: actions from a device buffer, packed observations only (obs = NULL)
            _lib.check(L.dq_policy_random_legal_next(h, p_legal, p_act, cur()))
            _lib.check(L.dq_env_step(h, p_act, None, p_rew, p_done, p_life, p_legal, 1, cur()))
        t_env = timed(ext_step, 100)
        t_act32 = timed(lambda i: act_iter(i, False), 40)
        agent.act_precision = "bf16"
        t_act = timed(lambda i: act_iter(i, False), 60)

        def train_iter(i):
            act_iter(i, True)
            if rring.filled >= 1:
                agent.train_on_ring(rring, i)
        t_train = timed(train_iter, 40)
        # the update alone (replay sample + 3 forwards + backward + Adam, batch 4096), in the three precision settings
        t_upd32 = timed(lambda i: agent.train_on_ring(rring, i), 30)
        upd = {"fp32_ms": t_upd32 * 1e3}
        try:
            agent_tc = A.DQNAgent(model=spec, nb_actions=env.num_actions, memory=A.SequentialMemory(limit=64 * n), nb_steps_warmup=0,
                                  target_model_update=10 ** 9, policy=A.EpsGreedyQPolicy(eps=0.1), test_policy=A.GreedyQPolicy(masked_greedy=True),
                                  enable_dueling_network=True, batch_size=4096, seed=SEED, device=dev, act_precision="bf16",
                                  target_precision="bf16", train_precision="fp32")
            agent_tc.compile(A.Adam(lr=1e-5), max_envs=n)
            upd["bf16_targets_fp32_gradients_ms"] = timed(lambda i: agent_tc.train_on_ring(rring, i), 30) * 1e3
            agent_tc.train_precision = "bf16"
            upd["bf16_ms"] = timed(lambda i: agent_tc.train_on_ring(rring, i), 30) * 1e3
            agent_tc.act_precision = "bf16"

            def train_iter_tc(i):
                rring.push_obs(rows_view)
                a = agent_tc._act(env, rows_ptr.value, i, 0.1, True)
                _lib.check(L.dq_env_step(h, vp(a), None, p_rew, p_done, p_life, p_legal, 1, cur()))
                rring.push_outcome(a, env.reward, env.done)
                agent_tc.train_on_ring(rring, i)
            t_train_tc = timed(train_iter_tc, 40)
            upd["train_bf16_env_steps_per_s"] = n / t_train_tc
            upd["train_bf16_ms_per_iteration"] = t_train_tc * 1e3
            upd["flops_per_update"] = 5 * 4096 * agent.model.flops_per_sample        # 3 forwards + backward (2 forward-equivalents)
            upd["bf16_tflops"] = upd["flops_per_update"] / (upd["bf16_ms"] * 1e-3) / 1e12
        except Exception as ex:      # noqa: BLE001 -- reported in the line
            upd["error"] = "%s: %s" % (type(ex).__name__, str(ex)[:300])
        flops = agent.model.flops_per_sample
        tf_peak = 1393.5
        try:
            tf_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"])
        except Exception:
            pass
        dqn = {"act_env_steps_per_s": n / t_act, "act_ms_per_iteration": t_act * 1e3,
               "train_env_steps_per_s": n / t_train, "train_ms_per_iteration": t_train * 1e3,
               "train_batch": 4096, "updates_per_iteration": 1, "update_ms": upd,
               "act_fp32_env_steps_per_s": n / t_act32,
               "policy_kernel_plus_env_step_external_actions_us": t_env * 1e6,
               "qnet_forward_fp32_ms": t_fwd * 1e3, "qnet_forward_fp32_tflops": flops * n / t_fwd / 1e12,
               "qnet_forward_bf16_ms": t_fwd_tc * 1e3, "qnet_forward_bf16_tflops": flops * n / t_fwd_tc / 1e12,
               "qnet_frac_of_bf16_sustained_peak": flops * n / t_fwd_tc / 1e12 / tf_peak,
               "cpu_serial_loop": cpu_serial_dqn_loop(min(6.0, max(0.5, args.cpu_seconds / 2))),
               "qnet_precision": "acting: bf16 tcgen05 (fp32 accumulate in TMEM); train_*: fp32 SIMT updates (Keras arithmetic); update_ms.bf16_ms / train_bf16_*: tcgen05 forward and backward with fp32 master weights (DQNAgent(train_precision='bf16'))",
               "qnet_flops_per_sample": flops, "policy": "eps-greedy 0.1 over legal actions, masked greedy"}

    if rank == 0:
        full, rest = divmod(K, ROLL)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u64", "data": "synthetic",
                "config": bench_config(world),
                "launch_detail": "the %d timed steps ran as %d dq_env_rollout_random launch(es): %d of %d steps%s" % (
                    K, timed_launches, full, ROLL, (" and one of %d" % rest) if rest else ""),
                "clocks": clocks, "e2e": e2e, "e2e_packed": e2e_packed,
                "gpu_launches": timed_launches, "single_step_launches": single,
                "roofline": roof, "roofline_scaling": scaling, "other_workloads": others, "cpu_baseline": cb, "dqn_dp": dqn_dp, "dqn": dqn,
                "logical_error_rate": (dqn_dp or {}).get("logical_error_rate")}
        print(json.dumps(line), flush=True)
    env.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def select_workload(name):
    global D, VD, P, MODEL, N_PER_GPU, BYTES_PER_STEP, WORKLOAD, METRIC
    global OBS_BYTES
    D, VD, P, MODEL, N_PER_GPU, BYTES_PER_STEP = WORKLOADS[name]
    OBS_BYTES = (VD + (2 if MODEL == "DP" else 1)) * (2 * D + 1) ** 2
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
    ap.add_argument("--no-experiments", action="store_true", help=argparse.SUPPRESS)      # accepted for older command lines; there is no such leg any more
    args = ap.parse_args()
    select_workload(args.workload)
    if args.workload != "c3":
        args.no_dqn = True
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
