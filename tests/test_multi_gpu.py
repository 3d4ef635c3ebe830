"""Tests of the data-parallel training exchange (csrc/dq_comm.cu).  The two-rank tests run whenever two GPUs are visible
(`gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`); the world-1 test runs the same kernel on a single-GPU box.
The host-side sharding logic is covered on CPU (gloo, world_size 2) in test_parallel_cpu.py."""
import ctypes as C
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _need_two_gpus():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")


def test_fused_allreduce_adam_with_itself_as_only_peer_equals_adam_step():
    """world = 1 through the C ABI: the exchange kernel signals / waits on its own flag slot, sums the one gradient and applies Adam --
    the same bits as dq_adam_step with grad_scale 1, over several updates (both gradient buffers, the tail that is not a multiple of 4)."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from deepq_decoding_b200 import _lib, agents as A
    from deepq_decoding_b200.qnet import device_view
    L = _lib.lib()
    dev = torch.device("cuda", 0)
    n = 193283
    gen = torch.Generator(device="cpu").manual_seed(3)
    p0 = torch.randn(n, generator=gen).to(dev)
    opt = A.Adam(lr=1e-3)
    pa, ma, va = p0.clone(), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    pb, mb, vb = p0.clone(), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    h = C.c_void_p()
    _lib.check(L.dq_comm_create(C.byref(h), 0, 1, n, 0))
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())
    launches0 = _lib.launch_count()
    for t in range(1, 6):
        g = torch.randn(n, generator=gen).to(dev)
        _lib.check(L.dq_adam_step(p(pa), p(ma), p(va), p(g), n, opt.lr, opt.beta_1, opt.beta_2, opt.epsilon, t, 1.0, st))
        gp = C.c_void_p()
        _lib.check(L.dq_comm_next_grads(h, C.byref(gp)))
        device_view(gp.value, (n,), "<f4", dev).copy_(g)
        _lib.check(L.dq_comm_allreduce_adam(h, p(pb), p(mb), p(vb), opt.lr, opt.beta_1, opt.beta_2, opt.epsilon, t, st))
    torch.cuda.synchronize()
    flag = C.c_int(1)
    _lib.check(L.dq_comm_status(h, C.byref(flag)))
    assert flag.value == 0, "no wait may time out"
    assert _lib.launch_count() - launches0 == 10
    assert torch.equal(pa, pb) and torch.equal(ma, mb) and torch.equal(va, vb)
    assert float((pb - p0).abs().max()) > 1e-3
    _lib.check(L.dq_comm_destroy(h))


def _kernel_worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from deepq_decoding_b200 import _lib, agents as A, parallel
    parallel.init("nccl")
    dev = torch.device("cuda", rank)
    L = _lib.lib()
    n = 193283                                           # the d=5 DP network: not a multiple of 4
    gen = torch.Generator(device="cpu").manual_seed(0)
    p0 = torch.randn(n, generator=gen).to(dev)
    opt = A.Adam(lr=1e-3)
    pa, ma, va = p0.clone(), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    pb, mb, vb = p0.clone(), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    comm = parallel.FusedAllreduceAdam(n, dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())
    for t in range(1, 8):
        g = torch.randn(n, generator=torch.Generator(device="cpu").manual_seed(1000 * t + rank)).to(dev)
        # baseline: NCCL all-reduce, then the stand-alone Adam kernel
        ga = g.clone()
        dist.all_reduce(ga)
        _lib.check(L.dq_adam_step(p(pa), p(ma), p(va), p(ga), n, opt.lr, opt.beta_1, opt.beta_2, opt.epsilon, t, 1.0 / world, st))
        # product: one kernel over peer memory
        comm.grads().copy_(g)
        comm.step(pb, mb, vb, opt, t, st)
    torch.cuda.synchronize()
    comm.check()
    gathered = [torch.zeros_like(pb) for _ in range(world)]
    dist.all_gather(gathered, pb)
    torch.save(dict(same_as_nccl=bool(torch.equal(pa, pb) and torch.equal(ma, mb) and torch.equal(va, vb)),
                    max_diff=float((pa - pb).abs().max()), identical_across_ranks=all(bool(torch.equal(x, gathered[0])) for x in gathered),
                    moved=float((pb - p0).abs().max())), os.path.join(out, "k%d.pt" % rank))
    dist.barrier()
    comm.close()
    dist.destroy_process_group()


def test_fused_allreduce_adam_is_bit_identical_to_nccl_then_adam(tmp_path):
    _need_two_gpus()
    mp.spawn(_kernel_worker, args=(2, 29500 + os.getpid() % 2000, str(tmp_path)), nprocs=2, join=True)
    for r in (0, 1):
        res = torch.load(os.path.join(tmp_path, "k%d.pt" % r), weights_only=False)
        assert res["moved"] > 1e-3
        assert res["identical_across_ranks"]
        assert res["same_as_nccl"], res["max_diff"]          # world 2: one fp32 add either way -> the same bits


def _fit_worker(rank, world, port, out, collective, total, train_precision="fp32"):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from deepq_decoding_b200 import agents as A, parallel
    from deepq_decoding_b200.envs import VecSurfaceCodeEnv
    parallel.init("nccl")
    dev = torch.device("cuda", rank)
    base, count = parallel.shard(total, rank, world)
    env = VecSurfaceCodeEnv(5, 0.007, 0.007, "X", False, 5, None, n_envs=count, seed=7, env_id_base=base, device=dev)
    spec = A.build_convolutional_nn([[64, 3, 2], [32, 2, 1], [32, 2, 1]], [[512, 0.2]], env.observation_space.shape, env.num_actions)
    pol = A.LinearAnnealedPolicy(A.EpsGreedyQPolicy(masked_greedy=False), attr="eps", value_max=1.0, value_min=0.02, value_test=0.0, nb_steps=40000)
    dqn = A.DQNAgent(model=spec, nb_actions=env.num_actions, memory=A.SequentialMemory(limit=100000), nb_steps_warmup=10000,
                     target_model_update=20000, policy=pol, test_policy=A.GreedyQPolicy(masked_greedy=True), gamma=0.99,
                     enable_dueling_network=True, batch_size=256, seed=0, device=dev, process_group=dist.group.WORLD, collective=collective,
                     act_precision="bf16" if train_precision == "bf16" else "fp32", target_precision=train_precision, train_precision=train_precision)
    dqn.compile(A.Adam(lr=1e-4), max_envs=count)
    parallel.broadcast_params_(dqn.model.params)
    dqn.target_params.copy_(dqn.model.params)
    start = dqn.model.params.clone()
    dqn.fit(env, nb_steps=60 * count, verbose=0, episode_averaging_length=500, success_threshold=1e9, stopping_patience=1e12)
    torch.cuda.synchronize()
    gathered = [torch.zeros_like(dqn.model.params) for _ in range(world)]
    dist.all_gather(gathered, dqn.model.params)
    torch.save(dict(identical=all(bool(torch.equal(x, gathered[0])) for x in gathered), updates=dqn.updates,
                    moved=float((dqn.model.params - start).abs().max()), fused=dqn.comm is not None,
                    finite=bool(torch.isfinite(dqn.model.params).all())), os.path.join(out, "f%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("collective,total", [("fused", 2048), ("nccl", 2048), ("fused", 2049), ("nccl", 2049)])
def test_sharded_fit_keeps_ranks_in_step(tmp_path, collective, total):
    """Also with UNEQUAL shards (2049 lattices = 1025 + 1024): every rank counts the largest shard's transitions per iteration, so
    warm-up, train_interval and termination fall on the same iteration everywhere and the number of collectives matches."""
    _need_two_gpus()
    mp.spawn(_fit_worker, args=(2, 31500 + os.getpid() % 2000, str(tmp_path), collective, total), nprocs=2, join=True)
    res = [torch.load(os.path.join(tmp_path, "f%d.pt" % r), weights_only=False) for r in (0, 1)]
    for r in res:
        assert r["identical"] and r["finite"] and r["updates"] > 20 and r["moved"] > 0 and r["fused"] == (collective == "fused")
    assert res[0]["updates"] == res[1]["updates"]


@pytest.mark.parametrize("collective", ["fused", "nccl"])
def test_sharded_fit_with_tensor_core_updates(tmp_path, collective):
    """The same sharded fit with bf16 updates: dq_qnet_backward_tc adds its gradient (fp32 atomics) straight into the exchange region
    (fused) or into the buffer NCCL reduces; the ranks' parameters stay bit-identical."""
    _need_two_gpus()
    mp.spawn(_fit_worker, args=(2, 33500 + os.getpid() % 2000, str(tmp_path), collective, 2048, "bf16"), nprocs=2, join=True)
    res = [torch.load(os.path.join(tmp_path, "f%d.pt" % r), weights_only=False) for r in (0, 1)]
    for r in res:
        assert r["identical"] and r["finite"] and r["updates"] > 20 and r["moved"] > 0 and r["fused"] == (collective == "fused")
    assert res[0]["updates"] == res[1]["updates"]
