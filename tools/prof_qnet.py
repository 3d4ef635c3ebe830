"""Runs the Q-network forward (fp32 SIMT and bf16 tcgen05) a few times on 16 384 synthetic packed observations.
Meant to sit under ncu:  ncu --metrics gpu__time_duration.sum --csv --log-file out.csv python tools/prof_qnet.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepq_decoding_b200.qnet import QNetwork  # noqa

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
q = QNetwork([[64, 3, 2], [32, 2, 1], [32, 2, 1]], [[512, 0.2]], (7, 11, 11), 51, dueling=True, max_batch=n)
boards = (torch.rand((n, 7, 11, 11), device="cuda") < 0.12).to(torch.uint8)
packed = q.pack(boards)
for _ in range(3):
    q.forward_packed(packed.data_ptr(), n, n)
for _ in range(3):
    q.forward_packed(packed.data_ptr(), n, n, precision="bf16")
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for prec in ("fp32", "bf16"):
    a.record()
    for _ in range(20):
        q.forward_packed(packed.data_ptr(), n, n, precision=prec)
    b.record()
    torch.cuda.synchronize()
    print(prec, "forward: %.1f us" % (a.elapsed_time(b) * 1e3 / 20))
