"""Iterative training over increasing error rates, in process.

The reference does this with a polled Slurm controller (cluster_scripts/d5_dp/Controller.py:15-33, 117-279; manuscript
tex:592-609): at error rate p_k it trains a grid of hyper-parameter configurations (one 4-core job each), scores every
configuration by its greedy test lifetime at p_k, keeps the best one if it beats the single-faulty-qubit threshold 1/p_k,
copies its `final_dqn_weights.h5f` -> `initial_dqn_weights.h5f` and `memory.p` forward, and moves to p_{k+1}.  Here a
"job" is a `DQNAgent.fit` call on vectorised lattices that takes seconds, so the grid is a loop; with several GPUs the
grid points can be dealt to ranks (`parallel.shard` over the grid) and the winner broadcast.
"""
import copy
import itertools
import os

import numpy as np

from . import agents as A
from .envs import VecSurfaceCodeEnv


def default_grid():
    """The reference's grid axes (Controller.py:28-33): learning rate x target-network period x final exploration."""
    return {"learning_rate": [1e-4, 5e-5, 1e-5], "target_network_update_freq": [2500, 5000], "final_eps": [0.02, 0.001]}


def train_point(spec_args, error_model, p, cfg, n_envs, steps, init=None, test_episodes=4096, seed=0, device="cuda:0", verbose=0):
    """One grid point at one error rate: build env + agent (optionally from carried-over weights/memory), fit, test.
    cfg keys follow the reference's config dicts: learning_rate, target_network_update_freq, final_eps, max_timesteps-like
    `steps`, exploration_fraction, learning_starts, buffer_size, batch_size, gamma."""
    cc, ff, d, vd = spec_args
    env = VecSurfaceCodeEnv(d, p, p, error_model, False, vd, None, n_envs=n_envs, seed=seed, device=device)
    spec = A.build_convolutional_nn(cc, ff, env.observation_space.shape, env.num_actions)
    eps0 = cfg.get("initial_eps", 1.0 if init is None else 0.3)
    policy = A.LinearAnnealedPolicy(A.EpsGreedyQPolicy(masked_greedy=False), attr="eps", value_max=eps0, value_min=cfg.get("final_eps", 0.02),
                                    value_test=0.0, nb_steps=cfg.get("exploration_fraction", 0.25) * steps)
    memory = A.SequentialMemory(limit=int(cfg.get("buffer_size", 2e6)), window_length=1)
    dqn = A.DQNAgent(model=spec, nb_actions=env.num_actions, memory=memory, nb_steps_warmup=int(cfg.get("learning_starts", 2e5)),
                     target_model_update=int(cfg.get("target_network_update_freq", 5000) * cfg.get("target_scale", 40)), policy=policy,
                     test_policy=A.GreedyQPolicy(masked_greedy=True), gamma=cfg.get("gamma", 0.99), enable_dueling_network=True,
                     batch_size=int(cfg.get("batch_size", 1024)), updates_per_step=int(cfg.get("updates_per_step", 2)), seed=seed,
                     device=device, act_precision=cfg.get("act_precision", "bf16"), target_precision=cfg.get("target_precision", "fp32"),
                     train_precision=cfg.get("train_precision", "fp32"))
    dqn.compile(A.Adam(lr=cfg.get("learning_rate", 1e-4)), max_envs=max(n_envs, test_episodes))
    if init is not None:
        dqn.model.params.copy_(init["params"].to(dqn.model.device))
        dqn.model.params_changed()
        dqn.target_params.copy_(dqn.model.params)
        if init.get("memory") is not None:
            dqn.load_memory(init["memory"], env)
    hist = dqn.fit(env, nb_steps=int(steps), verbose=verbose, episode_averaging_length=cfg.get("rolling_average_length", 5000),
                   success_threshold=cfg.get("success_threshold", 1e9), stopping_patience=cfg.get("stopping_patience", 1e12),
                   min_nb_steps=0).history
    test_env = VecSurfaceCodeEnv(d, p, p, error_model, False, vd, None, n_envs=test_episodes, seed=seed + 7919, device=device)
    life = np.array(dqn.test(test_env, nb_episodes=test_episodes, verbose=0).history["episode_lifetime"], dtype=np.float64)
    result = {"p_phys": p, "config": dict(cfg), "test_mean_lifetime": float(life.mean()), "test_se": float(life.std() / np.sqrt(len(life))),
              "threshold": 1.0 / p, "final_rolling_lifetime": hist["episode_lifetimes_rolling_avg"][-1] if hist.get("episode") else None}
    carry = {"params": dqn.model.params.detach().clone(), "memory": dqn.save_memory()}
    env.close(); test_env.close()
    return result, carry, dqn


def deal(points, rank, world):
    """Grid points of one error rate dealt round-robin to the ranks: [(grid index, config), ...] for `rank`.
    (The reference submits one Slurm job per point, Controller.py:161-279; here a rank is a GPU of the same box.)"""
    return [(gi, cfg) for gi, cfg in enumerate(points) if gi % world == rank]


def pick_winner(scores):
    """scores[r] = best greedy test lifetime rank r found at this error rate (None = it trained nothing).
    Returns the rank whose candidate wins; ties go to the lowest rank so every rank computes the same answer."""
    best = None
    for r, s in enumerate(scores):
        if s is not None and (best is None or s > scores[best]):
            best = r
    return best


def broadcast_carry(carry, src, group=None, device="cpu"):
    """The winner's hand-over state (flat parameters + replay snapshot: a dict of tensors and ints) from rank `src` to all."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(group)
    meta = [None]
    if rank == src:
        mem = carry.get("memory") or {}
        meta[0] = {"params": tuple(carry["params"].shape),
                   "tensors": {k: (tuple(v.shape), v.dtype) for k, v in mem.items() if torch.is_tensor(v)},
                   "plain": {k: v for k, v in mem.items() if not torch.is_tensor(v)}, "has_memory": carry.get("memory") is not None}
    dist.broadcast_object_list(meta, src=src, group=group)
    m = meta[0]
    params = carry["params"].to(device) if rank == src else torch.empty(m["params"], dtype=torch.float32, device=device)
    dist.broadcast(params, src=src, group=group)
    out = {"params": params, "memory": None}
    if m["has_memory"]:
        mem = dict(m["plain"])
        for k, (shape, dt) in m["tensors"].items():
            buf = carry["memory"][k].to(device) if rank == src else torch.empty(shape, dtype=dt, device=device)
            dist.broadcast(buf, src=src, group=group)
            mem[k] = buf.cpu()
        out["memory"] = mem
    return out


def iterative_training(error_rates, grid=None, error_model="DP", d=5, volume_depth=5, cc_layers=((64, 3, 2), (32, 2, 1), (32, 2, 1)),
                       ff_layers=((512, 0.2),), n_envs=4096, steps_per_point=2e7, test_episodes=4096, out_dir=None, seed=0,
                       device="cuda:0", verbose=0, max_grid_points=None, process_group=None):
    """Controller.py's state machine as a function: returns the per-rate winners and the final carry (weights + memory).
    With `process_group` (one process per GPU) the grid points of every error rate are dealt to the ranks, the winner is
    agreed on by an all-gather of the scores and its weights + replay memory are broadcast before the next rate."""
    rank, world = 0, 1
    if process_group is not None:
        import torch.distributed as dist
        rank, world = dist.get_rank(process_group), dist.get_world_size(process_group)
    grid = grid or default_grid()
    keys = sorted(grid)
    points = [dict(zip(keys, vals)) for vals in itertools.product(*(grid[k] for k in keys))]
    if max_grid_points:
        points = points[:max_grid_points]
    spec_args = ([list(l) for l in cc_layers], [list(l) for l in ff_layers], d, volume_depth)
    carry, winners = None, []
    for p in error_rates:
        best = None
        for gi, cfg in deal(points, rank, world):
            res, c, dqn = train_point(spec_args, error_model, p, cfg, n_envs, steps_per_point, init=copy.copy(carry),
                                      test_episodes=test_episodes, seed=seed + gi, device=device, verbose=verbose)
            if best is None or res["test_mean_lifetime"] > best[0]["test_mean_lifetime"]:
                best = (res, c, dqn)
            else:
                dqn.model.close()
        if world > 1:
            import torch.distributed as dist
            scores = [None] * world
            dist.all_gather_object(scores, (best[0]["test_mean_lifetime"], best[0]) if best else None, group=process_group)
            win = pick_winner([s[0] if s else None for s in scores])
            res = scores[win][1]
            c = broadcast_carry(best[1] if rank == win else None, win, process_group, device)
            dqn = best[2] if best else None
            if rank != win and dqn is not None:
                dqn.model.close(); dqn = None
        else:
            res, c, dqn = best
        res["beats_threshold"] = res["test_mean_lifetime"] > res["threshold"]
        winners.append(res)
        if out_dir and dqn is not None:
            os.makedirs(os.path.join(out_dir, str(p)), exist_ok=True)
            dqn.save_weights(os.path.join(out_dir, str(p), "final_dqn_weights.h5f"))
        if dqn is not None:
            dqn.model.close()
        if not res["beats_threshold"]:
            break                   # the reference stops the curriculum when no configuration beats 1/p (Controller.py:117-156)
        carry = c
    return winners, carry
