"""A few double-DQN updates on synthetic transitions (for an ncu launch list of the training step)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepq_decoding_b200 import agents as A  # noqa

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
TP = sys.argv[2] if len(sys.argv) > 2 else "fp32"
TRAIN = sys.argv[3] if len(sys.argv) > 3 else "fp32"
spec = A.build_convolutional_nn([[64, 3, 2], [32, 2, 1], [32, 2, 1]], [[512, 0.2]], (7, 11, 11), 51)
dqn = A.DQNAgent(model=spec, nb_actions=51, memory=A.SequentialMemory(limit=100), enable_dueling_network=True, batch_size=B, target_precision=TP, train_precision=TRAIN)
dqn.compile(A.Adam(lr=1e-4), max_envs=B)
boards = lambda: (torch.rand((B, 7, 11, 11), device="cuda") < 0.12).to(torch.uint8)
s0, s1 = dqn.model.pack(boards()), dqn.model.pack(boards())
act = torch.randint(0, 51, (B,), device="cuda", dtype=torch.int32)
rew = (torch.rand(B, device="cuda") < 0.3).float()
term = (torch.rand(B, device="cuda") < 0.1).to(torch.uint8)
for _ in range(3):
    dqn.update(s0, s1, act, rew, term)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    dqn.update(s0, s1, act, rew, term)
b.record()
torch.cuda.synchronize()
print("update (batch %d, no-grad forwards %s, gradients %s): %.1f us" % (B, TP, TRAIN, a.elapsed_time(b) * 1e3 / 20))
