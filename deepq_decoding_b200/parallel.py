"""Multi-GPU plumbing: one process per GPU (torch.distributed), lattices sharded by rank.

The reference has no collectives at all (one Slurm job per hyper-parameter point).  Here the independent
lattices of one run are partitioned contiguously over the ranks; the Philox stream id of a lattice is its
GLOBAL index (`env_id_base = shard base`), so a sharded run draws exactly the noise of the unsharded one.
Acting and evaluation need no communication; training all-reduces the flat gradient buffer once per update
(193 283 fp32 = 773 KB for the d=5 DP network) and evaluation sums (lifetime, episode) totals at the end.
"""
import ctypes as C
import os

import torch
import torch.distributed as dist


def shard(n_total, rank, world):
    """Contiguous shard [base, base+count) of n_total lattices for `rank`."""
    q, r = divmod(int(n_total), int(world))
    count = q + (1 if rank < r else 0)
    base = rank * q + min(rank, r)
    return base, count


def init(backend=None):
    """Join the job described by RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT (torchrun); returns (rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            local = int(os.environ.get("LOCAL_RANK", "0"))
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world


def allreduce_mean_(flat, group=None):
    """In-place mean of a flat gradient buffer over the ranks (identical Adam update everywhere afterwards)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, group=group)
        flat.div_(dist.get_world_size(group))
    return flat


def broadcast_params_(flat, src=0, group=None):
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(flat, src=src, group=group)
    return flat


def reduce_lifetimes(lifetimes, group=None):
    """(sum of lifetimes, sum of squares, episodes) over all ranks -> mean, standard error, episodes."""
    t = torch.as_tensor(lifetimes, dtype=torch.float64)
    acc = torch.tensor([float(t.sum()), float((t * t).sum()), float(t.numel())], dtype=torch.float64)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        if dist.get_backend(group) == "nccl":
            acc = acc.cuda()
        dist.all_reduce(acc, group=group)
        acc = acc.cpu()
    s, s2, n = acc.tolist()
    mean = s / max(n, 1.0)
    var = max(s2 / max(n, 1.0) - mean * mean, 0.0)
    return mean, (var / max(n, 1.0)) ** 0.5, int(n)


class FusedAllreduceAdam:
    """Gradient mean over the ranks + Keras-2 Adam in ONE kernel per rank over NVLink peer memory (csrc/dq_comm.cu).

    Replaces `dist.all_reduce(grads)` followed by `dq_adam_step`.  The 64-byte CUDA-IPC handles of the per-rank exchange
    regions travel once through `torch.distributed` (plumbing); the data path never touches NCCL.  Usage per update:
    `g = comm.grads()` -> backward writes this update's gradient into `g` -> `comm.step(params, m, v, optimizer, t, stream)`.
    """

    def __init__(self, n_params, device, group=None):
        from . import _lib
        self.L = _lib.lib()
        self._lib = _lib
        self.device = torch.device(device)
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.n = int(n_params)
        self._views = {}
        self._h = C.c_void_p()
        _lib.check(self.L.dq_comm_create(C.byref(self._h), self.rank, self.world, self.n, self.device.index or 0))
        mine = C.create_string_buffer(64)
        _lib.check(self.L.dq_comm_handle(self._h, mine))
        everyone = [None] * self.world
        dist.all_gather_object(everyone, bytes(mine.raw), group=group)
        _lib.check(self.L.dq_comm_connect(self._h, C.create_string_buffer(b"".join(everyone), 64 * self.world)))
        dist.barrier(group=group)              # every rank has mapped every region before the first signal is written

    def grads(self):
        """float32[n] view of the exchange buffer the NEXT update's gradient must be written to."""
        from .qnet import device_view
        ptr = C.c_void_p()
        self._lib.check(self.L.dq_comm_next_grads(self._h, C.byref(ptr)))
        view = self._views.get(ptr.value)                 # two buffers (update parity): wrap each once
        if view is None:
            view = self._views[ptr.value] = device_view(ptr.value, (self.n,), "<f4", self.device)
        return view

    def step(self, params, m, v, optimizer, t, stream):
        p = lambda x: C.c_void_p(x.data_ptr())
        self._lib.check(self.L.dq_comm_allreduce_adam(self._h, p(params), p(m), p(v), optimizer.lr, optimizer.beta_1, optimizer.beta_2,
                                                      optimizer.epsilon, int(t), stream))

    def check(self):
        """Raises if any wait on a peer timed out (the ranks fell out of step); synchronises the device."""
        flag = C.c_int(0)
        self._lib.check(self.L.dq_comm_status(self._h, C.byref(flag)))
        if flag.value:
            raise RuntimeError("fused all-reduce: a peer rank did not arrive (ranks out of step)")

    def close(self):
        if self._h:
            self.L.dq_comm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
