"""The reference's Slurm controller (cluster_scripts/d5_dp/Controller.py) as ONE multi-GPU job: at every error rate the
hyper-parameter grid is dealt to the ranks (one process per GPU), the winner is agreed on and its weights + replay memory
are broadcast before the next rate.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/grid_multi_gpu.py
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepq_decoding_b200 import curriculum, parallel  # noqa

rank, world = parallel.init()
dev = "cuda:%d" % int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(dev)
rates = [float(x) for x in os.environ.get("DQ_RATES", "0.001,0.003").split(",")]
steps = float(os.environ.get("DQ_STEPS", "1.5e7"))
grid = {"learning_rate": [1e-4, 5e-5], "target_network_update_freq": [2500, 5000], "final_eps": [0.02]}
t0 = time.time()
winners, carry = curriculum.iterative_training(rates, grid=grid, error_model=os.environ.get("DQ_MODEL", "X"), n_envs=4096, steps_per_point=steps,
                                               test_episodes=2048, device=dev, process_group=dist.group.WORLD if world > 1 else None)
torch.cuda.synchronize()
chk = carry["params"].double().sum().item() if carry else float("nan")
allchk = [None] * world
if world > 1:
    dist.all_gather_object(allchk, chk)
else:
    allchk = [chk]
if rank == 0:
    print(json.dumps({"world": world, "rates": rates, "grid_points": 4, "steps_per_point": steps, "seconds": time.time() - t0,
                      "winners": winners, "carry_identical_across_ranks": all(c == allchk[0] for c in allchk)}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
