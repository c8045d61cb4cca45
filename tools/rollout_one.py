"""Dev probe for ncu: fused Driving rollout launches.  usage: rollout_one.py [n]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from posggym_baselines_b200.envs.driving import BatchedDriving
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
env = BatchedDriving()
st = env.sample_initial_state(n, seed=1)
live = torch.zeros(1, dtype=torch.int64, device="cuda")
for k in range(3):
    env.rollout(st, 0, 50, 0.95, seed=1, live_count=live)
torch.cuda.synchronize()
print("ok", int(live.item()))
