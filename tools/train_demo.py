"""Train a decoder from scratch with DQNAgent.fit on vectorised lattices, then evaluate its logical lifetime.

    python tools/train_demo.py --model X --p 0.007 --envs 4096 --steps 40000000 --out gpurun_out/train_x.json

Writes the learning curve (rolling lifetime vs env-steps, from the fork-compatible history) and the greedy test
lifetime / logical error rate with its standard error."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepq_decoding_b200 import agents as A  # noqa
from deepq_decoding_b200.envs import VecSurfaceCodeEnv  # noqa

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="X")
ap.add_argument("--p", type=float, default=0.007)
ap.add_argument("--envs", type=int, default=4096)
ap.add_argument("--steps", type=float, default=4e7)
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--updates", type=int, default=2)
ap.add_argument("--lr", type=float, default=1e-4)
ap.add_argument("--eps-steps", type=float, default=1e7)
ap.add_argument("--eps-min", type=float, default=0.02)
ap.add_argument("--target-update", dest="target", type=float, default=2e5)
ap.add_argument("--buffer", type=float, default=2e6)
ap.add_argument("--warmup", type=float, default=2e5)
ap.add_argument("--gamma", type=float, default=0.99)
ap.add_argument("--seed", type=int, default=0)
ap.add_argument("--act", default="bf16")
ap.add_argument("--target", dest="target_prec", default="fp32", help="precision of the no-grad forwards inside an update")
ap.add_argument("--train", dest="train_prec", default="fp32", help="precision of an update's forward / backward pass (bf16: tcgen05, fp32 master weights)")
ap.add_argument("--test-episodes", type=int, default=8192)
ap.add_argument("--curriculum", default="", help="p:steps,p:steps,... trained in order, carrying weights and replay memory over (tex:592-609)")
ap.add_argument("--eps-max-continue", type=float, default=0.3)
ap.add_argument("--out", default="gpurun_out/train_demo.json")
a = ap.parse_args()

env = VecSurfaceCodeEnv(5, a.p, a.p, a.model, False, 5, None, n_envs=a.envs, seed=a.seed)
spec = A.build_convolutional_nn([[64, 3, 2], [32, 2, 1], [32, 2, 1]], [[512, 0.2]], env.observation_space.shape, env.num_actions)
policy = A.LinearAnnealedPolicy(A.EpsGreedyQPolicy(masked_greedy=False), attr="eps", value_max=1.0, value_min=a.eps_min, value_test=0.0, nb_steps=a.eps_steps)
dqn = A.DQNAgent(model=spec, nb_actions=env.num_actions, memory=A.SequentialMemory(limit=int(a.buffer)), nb_steps_warmup=int(a.warmup),
                 target_model_update=int(a.target), policy=policy, test_policy=A.GreedyQPolicy(masked_greedy=True), gamma=a.gamma,
                 enable_dueling_network=True, batch_size=a.batch, updates_per_step=a.updates, seed=a.seed, act_precision=a.act, target_precision=a.target_prec, train_precision=a.train_prec)
dqn.compile(A.Adam(lr=a.lr), max_envs=max(a.envs, a.test_episodes))
phases = [(a.p, a.steps)] if not a.curriculum else [(float(x.split(":")[0]), float(x.split(":")[1])) for x in a.curriculum.split(",")]
t0 = time.time()
curve, total_steps, hist = [], 0, None
for ph, (p_ph, n_ph) in enumerate(phases):
    env.p_phys = p_ph; env.p_meas = p_ph
    if ph > 0:      # a continued run restarts its step counter and exploration schedule (Single_Point_Continue_Training_Script.py:109-136)
        dqn.step = 0
        dqn.policy = A.LinearAnnealedPolicy(A.EpsGreedyQPolicy(masked_greedy=False), attr="eps", value_max=a.eps_max_continue,
                                            value_min=a.eps_min, value_test=0.0, nb_steps=min(a.eps_steps, n_ph / 3))
    hist = dqn.fit(env, nb_steps=int(n_ph), verbose=1, log_interval=4e6, episode_averaging_length=5000, success_threshold=1e9,
                   stopping_patience=1e12, min_nb_steps=0).history
    steps = np.array(hist["nb_steps"]); roll = np.array(hist["episode_lifetimes_rolling_avg"])
    idx = np.unique(np.linspace(0, len(steps) - 1, 40).astype(int))
    curve += [{"p_phys": p_ph, "env_steps": int(total_steps + steps[i]), "rolling_lifetime": float(roll[i])} for i in idx]
    total_steps += int(dqn.step)
t_train = time.time() - t0
a.p = phases[-1][0]
test_env = VecSurfaceCodeEnv(5, a.p, a.p, a.model, False, 5, None, n_envs=a.test_episodes, seed=a.seed + 1000)
th = dqn.test(test_env, nb_episodes=a.test_episodes, verbose=1).history
life = np.array(th["episode_lifetime"], dtype=np.float64)
res = {"config": vars(a), "train_seconds": t_train, "env_steps": int(total_steps), "updates": int(dqn.updates),
       "train_env_steps_per_s": total_steps / t_train, "episodes_last_phase": len(steps), "curve": curve,
       "test_mean_lifetime": float(life.mean()), "test_se": float(life.std() / np.sqrt(len(life))),
       "test_logical_error_rate_per_cycle": float(1.0 / life.mean()), "single_qubit_lifetime_1_over_p": 1.0 / a.p,
       "final_loss": float([x for x in hist["loss"] if x == x][-1]) if any(x == x for x in hist["loss"]) else None}
os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
json.dump(res, open(a.out, "w"), indent=1)
print(json.dumps({k: v for k, v in res.items() if k not in ("curve", "config")}))
