#!/usr/bin/env python
"""The reference's single-point training run (cluster_scripts/d5_dp/0.007/Single_Point_Continue_Training_Script.py with
trained_models/d5_dp/0.007/variable_config_84.p + fixed_config.p) through this package's drop-in classes, next to the history the
reference ships for that very run (trained_models/d5_dp/0.007/training_history.json, carried in tests/golden/history_pins.npz).

    python tools/train_reference_shape.py --envs 1            # the reference's loop shape: one lattice, one batch-32 update per env step
    python tools/train_reference_shape.py --envs 64 --batch 256 --updates 8
    python tools/train_reference_shape.py --envs 4096 --batch 4096 --updates 32 --steps 2.1e6

Hyper-parameters (the run's own): lr 1e-5, gamma 0.99, eps 0.5 -> 0.001 over 100 000 steps, warm-up 1000, target copy every 5000 steps,
buffer 50 000, rolling window 1000, patience 1000, min_nb_steps 100 000, max 1e6 steps, d=5 DP p=0.007 volume depth 5, masked_greedy False.
Like every run of the reference's curriculum above p=0.001 it CONTINUES from the best agent of the previous rate (Controller.py:187-270
copies final_dqn_weights.h5f -> initial_dqn_weights.h5f and memory.p): the shipped d5_dp/0.005 agent is the starting point here (fixture
tests/golden/dqn_d5_dp_0.005.npz); the replay memory of that run is not shipped, so this run starts with an empty one.

With N lattices per iteration the replay ratio of the reference (32 sampled transitions per env transition) is kept by
batch x updates_per_step = 32 N; the counters (warm-up, eps schedule, target period, nb_steps) stay in env transitions.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepq_decoding_b200 import agents as A  # noqa
from deepq_decoding_b200.envs import VecSurfaceCodeEnv  # noqa

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=1)
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--updates", type=int, default=1)
ap.add_argument("--steps", type=float, default=1e6)
ap.add_argument("--seed", type=int, default=0)
ap.add_argument("--init", default="0.005", help="shipped d5_dp agent to continue from ('none' = Glorot init)")
ap.add_argument("--test-episodes", type=int, default=4096)
ap.add_argument("--out", default="")
a = ap.parse_args()

cfg = dict(p=0.007, lr=1e-5, gamma=0.99, max_eps=0.5, final_eps=0.001, exploration_fraction=100000, learning_starts=1000,
           target_network_update_freq=5000, buffer_size=50000, rolling_average_length=1000, stopping_patience=1000, success_threshold=100000)
N = a.envs
env = VecSurfaceCodeEnv(5, cfg["p"], cfg["p"], "DP", False, 5, None, n_envs=N, seed=a.seed)
spec = A.build_convolutional_nn([[64, 3, 2], [32, 2, 1], [32, 2, 1]], [[512, 0.2]], env.observation_space.shape, env.num_actions)
policy = A.LinearAnnealedPolicy(A.EpsGreedyQPolicy(masked_greedy=False), attr="eps", value_max=cfg["max_eps"], value_min=cfg["final_eps"],
                                value_test=0.0, nb_steps=cfg["exploration_fraction"])
dqn = A.DQNAgent(model=spec, nb_actions=env.num_actions, memory=A.SequentialMemory(limit=max(cfg["buffer_size"], 4 * N), window_length=1),
                 nb_steps_warmup=cfg["learning_starts"], target_model_update=cfg["target_network_update_freq"], policy=policy,
                 test_policy=A.GreedyQPolicy(masked_greedy=True), gamma=cfg["gamma"], enable_dueling_network=True,
                 batch_size=a.batch, updates_per_step=a.updates, seed=a.seed, flush_interval=max(1, min(256, 65536 // N)))
dqn.compile(A.Adam(lr=cfg["lr"]), max_envs=max(N, a.test_episodes))
if a.init != "none":
    z = np.load(os.path.join(ROOT, "tests", "golden", "dqn_d5_dp_%s.npz" % a.init))
    dqn.model.set_keras_weights([(z["conv%d_k" % i], z["conv%d_b" % i]) for i in range(3)], [(z["dense%d_k" % i], z["dense%d_b" % i]) for i in range(3)])
    dqn.target_params.copy_(dqn.model.params)
t0 = time.time()
hist = dqn.fit(env, nb_steps=int(a.steps), action_repetition=1, callbacks=[], verbose=1, visualize=False, nb_max_start_steps=0,
               start_step_policy=None, log_interval=50000, nb_max_episode_steps=None, episode_averaging_length=cfg["rolling_average_length"],
               success_threshold=cfg["success_threshold"], stopping_patience=cfg["stopping_patience"], min_nb_steps=cfg["exploration_fraction"],
               single_cycle=False).history
secs = time.time() - t0
steps, roll = np.array(hist["nb_steps"]), np.array(hist["episode_lifetimes_rolling_avg"])
# the shipped run, sampled at the same step counts
pins = np.load(os.path.join(ROOT, "tests", "golden", "history_pins.npz"))
rs, rr = pins["d5_dp_0.007/nb_steps"], pins["d5_dp_0.007/rolling"]
marks = [m for m in (2000, 5000, 10000, 20000, 50000, 100000, 150000, 200000, 300000, 400000, 500000, 522725, 750000, 1000000, 2000000) if m <= steps[-1]]
at = lambda s_, r_, m: float(r_[min(np.searchsorted(s_, m), len(r_) - 1)])
curve = [{"env_steps": m, "rolling_lifetime": at(steps, roll, m), "reference_rolling_lifetime": at(rs, rr, m) if m <= rs[-1] else None} for m in marks]
test_env = VecSurfaceCodeEnv(5, cfg["p"], cfg["p"], "DP", False, 5, None, n_envs=a.test_episodes, seed=a.seed + 1000)
life = np.array(dqn.test(test_env, nb_episodes=a.test_episodes, verbose=0).history["episode_lifetime"], dtype=np.float64)
first_eps = next((x for x in hist["mean_eps"] if x == x), None)
res = {"what": "reference single-point run d5_dp/0.007 (variable_config_84.p) through DQNAgent.fit, continuing from the shipped d5_dp/%s agent" % a.init,
       "lattices": N, "batch": a.batch, "updates_per_iteration": a.updates, "replay_ratio_samples_per_transition": a.batch * a.updates / N,
       "hyper_parameters": cfg, "env_steps": int(steps[-1]), "episodes": int(len(steps)), "updates": int(dqn.updates), "seconds": secs,
       "env_steps_per_s": float(steps[-1] / secs), "stopped_improving": bool(hist["stopped_improving"][-1]),
       "best_rolling_avg": float(hist["best_rolling_avg"][-1]), "final_rolling_avg": float(roll[-1]),
       "first_logged_mean_eps": first_eps, "curve": curve,
       "reference_run": {"env_steps": int(rs[-1]), "episodes": int(len(rs)), "final_rolling_avg": float(rr[-1]), "best_rolling_avg": float(pins["d5_dp_0.007/best_rolling"][-1]),
                         "stopped_improving": bool(pins["d5_dp_0.007/stopped_improving"][-1]), "env_steps_per_s_logged": 41.6},
       "greedy_test": {"episodes": int(len(life)), "mean_lifetime": float(life.mean()), "standard_error": float(life.std() / np.sqrt(len(life))),
                       "reference_published_for_its_run": 270.42}}
out = a.out or os.path.join(ROOT, "gpurun_out", "train_reference_shape_n%d.json" % N)
os.makedirs(os.path.dirname(out), exist_ok=True)
json.dump(res, open(out, "w"), indent=1)
print(json.dumps({k: v for k, v in res.items() if k not in ("hyper_parameters",)}))
