"""Dev probe: per-step time of the captured collector loop (configs[4]: LBF-v3, 8 192 envs per GPU, 64 steps)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from posggym_baselines_b200.envs.lbf import LBFModel
from posggym_baselines_b200.ppo.collector import DeviceCollector
from tests.test_policy_gpu import RefNet

dev = torch.device("cuda")
net = RefNet(300, 6).to(dev)
for n_envs in (8192, 65536):
    col = DeviceCollector(LBFModel(device=dev), net.state_dict(), n_envs, rollout_length=64, seed=1)
    col.rollout_eager(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); col.rollout_eager(); e1.record(); torch.cuda.synchronize()
    eager = e0.elapsed_time(e1) / 64
    col.capture(); col.rollout(); torch.cuda.synchronize()
    e0.record()
    for _ in range(5): col.rollout()
    e1.record(); torch.cuda.synchronize()
    graph = e0.elapsed_time(e1) / 5 / 64
    print(f"envs {n_envs}: eager {1e3*eager:.1f} us/step, graph {1e3*graph:.1f} us/step, {n_envs/(graph*1e-3):.3e} env-steps/s (graph)")
