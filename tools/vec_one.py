"""Dev probe for ncu: a few posg_vec_step launches of one env at one size.  usage: vec_one.py lbf 1048576"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from posggym_baselines_b200.envs.env import make
from posggym_baselines_b200.vector import BatchedVecEnv

name = {"lbf": "LevelBasedForaging-v3", "driving": "Driving-v1", "pp": "PredatorPrey-v0"}[sys.argv[1]]
nv = int(sys.argv[2])
model = make(name).model
ve = BatchedVecEnv(model, nv, seed=1)
ve.reset()
nact = model.action_spaces["0"].n
acts = torch.randint(0, nact, (6, nv * 2), device="cuda", dtype=torch.int64)
for k in range(6):
    ve._global_step += 1
    ve._launch(acts[k], 0)
torch.cuda.synchronize()
print("ok")
