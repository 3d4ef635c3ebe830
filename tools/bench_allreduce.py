"""Gradient exchange of one training update on N GPUs: NCCL all-reduce + Adam kernel vs the fused peer-memory kernel.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_allreduce.py

Timed on the device (CUDA events on the launching stream, max over ranks), 300 updates after 30 warm-ups."""
import ctypes as C
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepq_decoding_b200 import _lib, agents as A, parallel  # noqa

rank, world = parallel.init("nccl")
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
torch.cuda.set_device(dev)
L = _lib.lib()
n = 193283
opt = A.Adam(lr=1e-4)
params, m, v = torch.randn(n, device=dev), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
g = torch.randn(n, device=dev)
comm = parallel.FusedAllreduceAdam(n, dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
p = lambda t: C.c_void_p(t.data_ptr())
t = 0


def nccl_update():
    global t
    t += 1
    dist.all_reduce(g)
    _lib.check(L.dq_adam_step(p(params), p(m), p(v), p(g), n, opt.lr, opt.beta_1, opt.beta_2, opt.epsilon, t, 1.0 / world, st))


def fused_update():
    global t
    t += 1
    comm.grads()                   # backward would write here; the exchange cost does not depend on the values
    comm.step(params, m, v, opt, t, st)


def timed(fn, iters=300, warm=30):
    for _ in range(warm):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms) * 1e3


res = {"world": world, "floats": n, "nccl_allreduce_then_adam_us": timed(nccl_update), "fused_peer_allreduce_adam_us": timed(fused_update)}
g.normal_()
res["nccl_allreduce_then_adam_us_2"] = timed(nccl_update)
res["fused_peer_allreduce_adam_us_2"] = timed(fused_update)
comm.check()
if rank == 0:
    print(json.dumps(res), flush=True)
dist.barrier()
comm.close()
dist.destroy_process_group()
