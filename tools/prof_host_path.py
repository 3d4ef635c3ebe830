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
