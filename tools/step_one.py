"""Dev probe for ncu: a few Driving model.step launches (random-action sweep).  usage: step_one.py [n]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from posggym_baselines_b200.envs.driving import BatchedDriving

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
env = BatchedDriving()
st = env.sample_initial_state(n, seed=1)
live = torch.zeros(1, dtype=torch.int64, device="cuda")
cur = st
outs = [None, None]
for k in range(8):
    outs[k & 1] = env.step(cur, seed=1, step_idx=k, out=outs[k & 1], live_count=live)
    cur = outs[k & 1].state
torch.cuda.synchronize()
print("ok", int(live.item()))
