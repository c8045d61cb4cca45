"""Dev probe for ncu: a few generic model.step launches.  usage: model_one.py lbf|pp [n]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from posggym_baselines_b200.envs.lbf import BatchedLBF
from posggym_baselines_b200.envs.predator_prey import BatchedPredatorPrey

env = BatchedLBF() if sys.argv[1] == "lbf" else BatchedPredatorPrey()
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 22
st = env.sample_initial_state(n, seed=1)
out = None
cur = st
for k in range(8):
    out = env.step(cur, seed=1, step_idx=k, out=out)
    cur = out.state
torch.cuda.synchronize()
print("ok")
