"""Dev probe for ncu: a few posg_vec_embed launches on LevelBasedForaging-v3 at 1 M envs."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from posggym_baselines_b200.vector import load_vec_env

ve = load_vec_env("LevelBasedForaging-v3", num_envs=1 << 20, seed=1)
ve.reset()
w = torch.randn(64, 300, device="cuda") * 0.1
b = torch.zeros(64, device="cuda")
for _ in range(4):
    ve.embed(w, b)
torch.cuda.synchronize()
print("ok")
