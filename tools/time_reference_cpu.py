#!/usr/bin/env python
"""Times the UNMODIFIED reference environment on this box's host cores (build container only: needs /root/reference).

    python tools/time_reference_cpu.py [--seconds 30] [--procs P] > profiles/reference_cpu_<box>.json

Workload = BASELINE config C3 per process: cluster_scripts/d5_dp/Environments.py as shipped (gym stubbed, see
oracle/ref_harness.py), d=5 depolarising p_phys=p_meas=0.007, volume depth 5, the reference's OWN random draws
(numpy global MT19937), the shipped referee MLP (example_notebooks/referee_decoders/nn_d5_DP_p5) evaluated per step,
batch 1, by torch-CPU fp32 (stand-in for Keras `predict`: keras / tensorflow are not installable here), uniform
random-legal policy.  P independent processes (default: every core we may use), each for `--seconds` of wall time;
reported: per-process and aggregate env-steps/s.  This is the number BASELINE.md section 2 asks for; bench.py prints the
committed result next to the C port it times live (the Python reference cannot travel to the GPU box).
"""
import argparse
import json
import multiprocessing as mp
import os
import platform
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker(idx, seconds, q):
    import numpy as np
    import torch
    from oracle import ref_harness as RH
    from deepq_decoding_b200 import referee as R
    torch.set_num_threads(1)
    E, FL = RH.reference_modules()
    layers = [(torch.from_numpy(np.asarray(W, np.float32)), torch.from_numpy(np.asarray(b, np.float32)))
              for W, b in R.load_keras_mlp("/root/reference/example_notebooks/referee_decoders/nn_d5_DP_p5")]

    class TorchReferee:                       # the duck-typed static_decoder of Environments.py:144
        def predict(self, x, batch_size=1, verbose=0):
            with torch.no_grad():
                h = torch.from_numpy(np.asarray(x, np.float32))
                for i, (W, b) in enumerate(layers):
                    h = h @ W + b
                    if i + 1 < len(layers):
                        h = torch.relu(h)
                return torch.softmax(h, dim=1).numpy()

    np.random.seed(1000 + idx)
    env = E.Surface_Code_Environment_Multi_Decoding_Cycles(d=5, p_phys=0.007, p_meas=0.007, error_model="DP", use_Y=False,
                                                           volume_depth=5, static_decoder=TorchReferee())
    env.reset()
    for _ in range(200):                       # warm-up
        _, _, done, _ = env.step(int(np.random.choice(sorted(env.legal_actions))))
        if done:
            env.reset()
    steps, episodes, t0 = 0, 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        for _ in range(50):
            _, _, done, _ = env.step(int(np.random.choice(sorted(env.legal_actions))))
            steps += 1
            if done:
                env.reset()
                episodes += 1
    q.put((idx, steps, episodes, time.perf_counter() - t0))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=30.0)
    ap.add_argument("--procs", type=int, default=0)
    args = ap.parse_args()
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        cores = os.cpu_count() or 1
    procs = args.procs or cores
    q = mp.Queue()
    ps = [mp.Process(target=worker, args=(i, args.seconds, q)) for i in range(procs)]
    t0 = time.perf_counter()
    for p in ps:
        p.start()
    res = sorted(q.get() for _ in ps)
    for p in ps:
        p.join()
    wall = time.perf_counter() - t0
    per = [s / el for _, s, _, el in res]
    cpu = ""
    try:
        cpu = [l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
    except Exception:
        pass
    print(json.dumps({
        "what": "UNMODIFIED reference Surface_Code_Environment_Multi_Decoding_Cycles (cluster_scripts/d5_dp/Environments.py), d=5 DP p=0.007 "
                "volume_depth=5, shipped referee MLP by torch-CPU fp32 (batch 1 per step), uniform random-legal policy, numpy's own RNG",
        "processes": procs, "cores_available": cores, "seconds_per_process": args.seconds, "wall_seconds": wall,
        "env_steps_per_s_aggregate": sum(per), "env_steps_per_s_per_process_mean": sum(per) / len(per),
        "env_steps_per_s_per_process_min": min(per), "env_steps_per_s_per_process_max": max(per),
        "episodes": sum(e for _, _, e, _ in res), "steps": sum(s for _, s, _, _ in res),
        "host": {"cpu": cpu, "machine": platform.machine(), "python": platform.python_version()},
        "unit": "env-steps/s"}))


if __name__ == "__main__":
    main()
