"""Data-parallel training across the GPUs of one box: lattices sharded by rank, one gradient all-reduce per update.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/train_multi_gpu.py

Checks that every rank holds bit-identical parameters after training (same all-reduced gradients, same Adam step)
and prints aggregate training throughput and the merged greedy test lifetime."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepq_decoding_b200 import agents as A, parallel  # noqa
from deepq_decoding_b200.envs import VecSurfaceCodeEnv  # noqa

rank, world = parallel.init()
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
torch.cuda.set_device(dev)
N_TOTAL, STEPS = 8192, float(os.environ.get("DQ_STEPS", "1.2e7"))
base, count = parallel.shard(N_TOTAL, rank, world)
env = VecSurfaceCodeEnv(5, 0.007, 0.007, "X", False, 5, None, n_envs=count, seed=7, env_id_base=base, device=dev)
spec = A.build_convolutional_nn([[64, 3, 2], [32, 2, 1], [32, 2, 1]], [[512, 0.2]], env.observation_space.shape, env.num_actions)
pol = A.LinearAnnealedPolicy(A.EpsGreedyQPolicy(masked_greedy=False), attr="eps", value_max=1.0, value_min=0.02, value_test=0.0, nb_steps=4e6 / world)
dqn = A.DQNAgent(model=spec, nb_actions=env.num_actions, memory=A.SequentialMemory(limit=int(2e6 / world)), nb_steps_warmup=int(1e5 / world),
                 target_model_update=int(2e5 / world), policy=pol, test_policy=A.GreedyQPolicy(masked_greedy=True), gamma=0.99,
                 enable_dueling_network=True, batch_size=1024 // world, updates_per_step=2, seed=0, device=dev, act_precision="bf16",
                 process_group=dist.group.WORLD if world > 1 else None)
dqn.compile(A.Adam(lr=1e-4), max_envs=max(count, 4096))
parallel.broadcast_params_(dqn.model.params)
dqn.target_params.copy_(dqn.model.params)
t0 = time.time()
dqn.fit(env, nb_steps=int(STEPS / world), verbose=0, episode_averaging_length=2000, success_threshold=1e9, stopping_patience=1e12)
torch.cuda.synchronize()
t_train = time.time() - t0
# every rank must hold the same parameters
chk = dqn.model.params.double().sum().reshape(1)
allchk = [torch.zeros_like(chk) for _ in range(world)]
if world > 1:
    dist.all_gather(allchk, chk)
else:
    allchk = [chk]
same = all(float(c) == float(allchk[0]) for c in allchk)
test_env = VecSurfaceCodeEnv(5, 0.007, 0.007, "X", False, 5, None, n_envs=4096, seed=99, env_id_base=100000 + rank * 4096, device=dev)
h = dqn.test(test_env, nb_episodes=4096, verbose=0).history
mean, se, n = parallel.reduce_lifetimes(h["episode_lifetime"])
if rank == 0:
    print(json.dumps({"world": world, "params_identical_across_ranks": same, "env_steps_total": int(dqn.step * world), "updates": dqn.updates,
                      "train_seconds": t_train, "train_env_steps_per_s": dqn.step * world / t_train, "test_episodes": n,
                      "test_mean_lifetime": mean, "test_se": se}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
