"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: sharding, gradient mean, evaluation merge."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deepq_decoding_b200 import parallel


def test_shards_partition_the_lattices():
    for n, world in ((131072, 8), (65536, 8), (1000, 3), (5, 8)):
        spans = [parallel.shard(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == n
        for (b0, c0), (b1, _) in zip(spans, spans[1:]):
            assert b0 + c0 == b1


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    r, w = parallel.init("gloo")
    assert (r, w) == (rank, world)
    g = torch.full((1000,), float(rank + 1))
    parallel.allreduce_mean_(g)
    p = torch.arange(10.0) * (rank + 1)
    parallel.broadcast_params_(p)
    rng = np.random.default_rng(rank)
    life = rng.integers(1, 100, size=500 + 100 * rank) * 5
    mean, se, n = parallel.reduce_lifetimes(life)
    torch.save(dict(g=g, p=p, mean=mean, se=se, n=n, life=life), os.path.join(out, "r%d.pt" % rank))
    dist.destroy_process_group()


def test_gloo_world2(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (torch.load(os.path.join(tmp_path, "r%d.pt" % r), weights_only=False) for r in (0, 1))
    assert torch.allclose(r0["g"], torch.full((1000,), 1.5)) and torch.equal(r0["g"], r1["g"])
    assert torch.equal(r0["p"], torch.arange(10.0)) and torch.equal(r1["p"], torch.arange(10.0))
    both = np.concatenate([r0["life"], r1["life"]])
    assert r0["n"] == r1["n"] == len(both)
    assert r0["mean"] == pytest.approx(both.mean()) and r1["mean"] == pytest.approx(both.mean())
    assert r0["se"] == pytest.approx(both.std() / np.sqrt(len(both)))


def test_grid_points_are_dealt_and_the_winner_is_agreed():
    from deepq_decoding_b200 import curriculum
    pts = [{"i": i} for i in range(7)]
    dealt = [curriculum.deal(pts, r, 3) for r in range(3)]
    assert sorted(gi for d in dealt for gi, _ in d) == list(range(7))
    assert [gi for gi, _ in dealt[1]] == [1, 4]
    assert curriculum.pick_winner([10.0, None, 30.0, 30.0]) == 2          # ties -> lowest rank, idle ranks skipped
    assert curriculum.pick_winner([None, None]) is None


def _carry_worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from deepq_decoding_b200 import curriculum
    parallel.init("gloo")
    carry = None
    if rank == 1:                                            # rank 1 holds the winning candidate
        carry = {"params": torch.arange(1000, dtype=torch.float32) * 0.5,
                 "memory": {"obs": torch.arange(24, dtype=torch.int64).reshape(2, 3, 4), "act": torch.ones(2, 4, dtype=torch.int32),
                            "head": 1, "filled": 2}}
    got = curriculum.broadcast_carry(carry, src=1, device="cpu")
    torch.save(got, os.path.join(out, "c%d.pt" % rank))
    dist.destroy_process_group()


def test_winner_state_reaches_every_rank(tmp_path):
    port = 27500 + (os.getpid() % 2000)
    mp.spawn(_carry_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in (0, 1):
        got = torch.load(os.path.join(tmp_path, "c%d.pt" % r), weights_only=False)
        assert torch.equal(got["params"], torch.arange(1000, dtype=torch.float32) * 0.5)
        assert torch.equal(got["memory"]["obs"], torch.arange(24, dtype=torch.int64).reshape(2, 3, 4))
        assert got["memory"]["act"].dtype == torch.int32 and got["memory"]["head"] == 1 and got["memory"]["filled"] == 2
