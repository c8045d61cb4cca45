"""Dev probe for ncu: the captured collector's kernels at 8 192 envs (eager, a few steps)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from posggym_baselines_b200.envs.lbf import LBFModel
from posggym_baselines_b200.ppo.collector import DeviceCollector
from tests.test_policy_gpu import RefNet
dev = torch.device("cuda")
col = DeviceCollector(LBFModel(device=dev), RefNet(300, 6).to(dev).state_dict(), int(sys.argv[1]) if len(sys.argv) > 1 else 8192,
                      rollout_length=8, seed=1)
col.rollout_eager(); col.rollout_eager()
torch.cuda.synchronize()
print("ok")
