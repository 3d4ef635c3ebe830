"""Where a host-buffer step spends its time (C3 workload): host policy, the split call with and without observations, the packed call,
the two-handle overlap pattern -- for a few sizes of the library's host thread pool (DQ_HOST_THREADS, read once per process: one child each)."""
import ctypes as C
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import numpy as np
    import torch
    from deepq_decoding_b200 import _lib
    from deepq_decoding_b200.envs import VecSurfaceCodeEnv
    L = _lib.lib()
    n = int(os.environ.get("DQ_N", "16384"))
    env = VecSurfaceCodeEnv(5, 0.007, 0.007, "DP", False, 5, None, n_envs=n, seed=3)
    hb = env._host_buffers()
    pk = env._packed_host_buffer()
    hp = lambda t: C.c_void_p(t.data_ptr())
    env.reset_host()
    res = {"threads": os.environ.get("DQ_HOST_THREADS", "default"), "lattices": n}

    def timed(fn, iters=40):
        for i in range(4):
            fn(i)
        t0 = time.perf_counter()
        for i in range(iters):
            fn(4 + i)
        return (time.perf_counter() - t0) / iters * 1e6
    h = env._h
    res["policy_us"] = timed(lambda i: L.dq_policy_random_legal_host(h, hp(hb["legal"]), i, hp(hb["actions"])))
    res["step_small_outputs_us"] = timed(lambda i: (L.dq_env_step_host_begin(h, hp(hb["actions"]), None, hp(hb["reward"]), hp(hb["done"]), hp(hb["lifetime"]), hp(hb["legal"]), 1),
                                                    L.dq_env_step_host_end(h)))
    res["step_with_obs_us"] = timed(lambda i: (L.dq_env_step_host_begin(h, hp(hb["actions"]), hp(hb["obs"]), hp(hb["reward"]), hp(hb["done"]), hp(hb["lifetime"]), hp(hb["legal"]), 1),
                                               L.dq_env_step_host_end(h)))
    res["step_packed_us"] = timed(lambda i: L.dq_env_step_host_packed(h, hp(hb["actions"]), hp(pk), hp(hb["reward"]), hp(hb["done"]), hp(hb["lifetime"]), hp(hb["legal"]), 1))
    res["policy_plus_step_with_obs_us"] = timed(lambda i: (L.dq_policy_random_legal_host(h, hp(hb["legal"]), i, hp(hb["actions"])),
                                                           L.dq_env_step_host_begin(h, hp(hb["actions"]), hp(hb["obs"]), hp(hb["reward"]), hp(hb["done"]), hp(hb["lifetime"]), hp(hb["legal"]), 1),
                                                           L.dq_env_step_host_end(h)))
    res["python_wrappers_step_us"] = timed(lambda i: (env.random_legal_actions_host(i), env.step_host_begin(), env.step_host_end()))
    # the two-handle overlap pattern, with the time spent in each kind of call
    half = n // 2
    envs = [VecSurfaceCodeEnv(5, 0.007, 0.007, "DP", False, 5, None, n_envs=half, seed=4, env_id_base=k * half) for k in (0, 1)]
    hbs = [e._host_buffers() for e in envs]
    for e in envs:
        e.reset_host()
    acc = {"policy": 0.0, "begin": 0.0, "end": 0.0}

    def begin(k, i, count):
        t0 = time.perf_counter()
        L.dq_policy_random_legal_host(envs[k]._h, hp(hbs[k]["legal"]), i, hp(hbs[k]["actions"]))
        t1 = time.perf_counter()
        rc = L.dq_env_step_host_begin(envs[k]._h, hp(hbs[k]["actions"]), hp(hbs[k]["obs"]), hp(hbs[k]["reward"]), hp(hbs[k]["done"]), hp(hbs[k]["lifetime"]), hp(hbs[k]["legal"]), 1)
        assert rc == 0
        t2 = time.perf_counter()
        if count:
            acc["policy"] += t1 - t0; acc["begin"] += t2 - t1

    def end(k, count):
        t0 = time.perf_counter()
        assert L.dq_env_step_host_end(envs[k]._h) == 0
        if count:
            acc["end"] += time.perf_counter() - t0
    for k in (0, 1):
        begin(k, 0, False)
    for i in range(4):
        for k in (0, 1):
            end(k, False); begin(k, 1 + i, False)
    iters = 40
    t0 = time.perf_counter()
    for i in range(iters):
        for k in (0, 1):
            end(k, True); begin(k, 10 + i, True)
    total = time.perf_counter() - t0
    for k in (0, 1):
        end(k, False)
    res["two_handles_us_per_full_step"] = total / iters * 1e6
    res["two_handles_breakdown_us"] = {k: v / iters * 1e6 for k, v in acc.items()}
    print("HOSTPROF " + json.dumps(res), flush=True)


if __name__ == "__main__":
    if os.environ.get("DQ_HOSTPROF_CHILD"):
        child()
    else:
        out = []
        for th, ch in (("default", "2"), ("default", "1"), ("default", "4"), ("8", "2"), ("4", "2")):
            envv = dict(os.environ, DQ_HOSTPROF_CHILD="1", DQ_HOST_CHUNKS=ch)
            if th != "default":
                envv["DQ_HOST_THREADS"] = th
            r = subprocess.run([sys.executable, os.path.abspath(__file__)], env=envv, capture_output=True, text=True, timeout=120)
            line = [l for l in r.stdout.splitlines() if l.startswith("HOSTPROF ")]
            res = json.loads(line[-1][9:]) if line else {"threads": th, "error": (r.stderr or r.stdout)[-300:]}
            res["chunks"] = ch
            out.append(res)
        print(json.dumps({"cores": os.cpu_count(), "runs": out}))
