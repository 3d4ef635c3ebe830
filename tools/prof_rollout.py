"""Single-step launches vs one multi-step rollout launch at the C3 workload (device time, CUDA events)."""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepq_decoding_b200 import _lib  # noqa
from deepq_decoding_b200.envs import VecSurfaceCodeEnv  # noqa

N = int(os.environ.get("DQ_N", "16384"))
D = int(os.environ.get("DQ_D", "5"))
env = VecSurfaceCodeEnv(D, 0.007 if D == 5 else 0.011, 0.007 if D == 5 else 0.011, "DP", False, D, None, n_envs=N, seed=3)
L = _lib.lib()
env.reset()
SLOTS = max(2, int(300e6 // env.obs.numel()) + 1)
ring = torch.zeros((SLOTS,) + tuple(env.obs.shape), dtype=torch.uint8, device="cuda")
p = lambda x: C.c_void_p(x.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
act = torch.zeros(N, dtype=torch.int32, device="cuda")
res = {"lattices": N, "d": D, "ring_slots": SLOTS}


def timed(fn, steps_per_call, calls):
    for _ in range(3):
        fn(0)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(calls):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / (calls * steps_per_call)


def single(i):
    _lib.check(L.dq_env_step_random(env._h, p(ring[i % SLOTS]), p(env.reward), p(env.done), p(env.lifetime), p(env.legal_mask), p(act), 1, st))


ONLY = int(os.environ.get("DQ_ONLY_ROLLOUT", "0"))      # ncu capture: just a few rollout launches of this many steps
if not ONLY:
    res["single_step_launches_us_per_step"] = timed(single, 1, 400)
for S in ((ONLY,) if ONLY else (4, 16, 64, 256)):
    rew = torch.empty((S, N), dtype=torch.float32, device="cuda")
    done = torch.empty((S, N), dtype=torch.uint8, device="cuda")
    life = torch.empty((S, N), dtype=torch.int32, device="cuda")
    legal = torch.empty((S, N, env.mask_words), dtype=torch.int64, device="cuda")
    acts = torch.empty((S, N), dtype=torch.int32, device="cuda")

    def roll(i, S=S):
        _lib.check(L.dq_env_rollout_random(env._h, S, p(ring), SLOTS, (i * S) % SLOTS, p(rew), p(done), p(life), p(legal), p(acts), 1, st))

    us = timed(roll, S, int(os.environ.get("DQ_CALLS", "2")) if ONLY else max(4, 1024 // S))     # DQ_CALLS: A/B runs time more launches than an ncu capture needs
    res["rollout_%d_us_per_step" % S] = us
    res["rollout_%d_env_steps_per_s" % S] = N / us * 1e6
print(json.dumps(res))
