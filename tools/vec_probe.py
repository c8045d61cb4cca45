"""Dev probe: time posg_vec_step for each env at several batch sizes (CUDA events)."""
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from posggym_baselines_b200.envs.driving import DrivingModel
from posggym_baselines_b200.envs.lbf import LBFModel
from posggym_baselines_b200.envs.predator_prey import PredatorPreyModel
from posggym_baselines_b200.vector import BatchedVecEnv

dev = torch.device("cuda", 0)
sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [8192, 65536, 1 << 20]
out = {}
for name, model, nact in (("lbf", LBFModel(device=dev), 6), ("driving", DrivingModel(device=dev), 5),
                          ("pp", PredatorPreyModel(device=dev), 5)):
    for nv in sizes:
        ve = BatchedVecEnv(model, nv, seed=1)
        ve.reset()
        acts = torch.randint(0, nact, (64, nv * 2), device=dev, dtype=torch.int64)
        for k in range(8):
            ve.step(acts[k])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(64):
            ve.step(acts[k])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 64
        bpe = 8 + 8 + 16 + 2 * ve.flat_dim * 4 + 8 + 2 + 2 + 1 + 8 + 4
        out[f"{name}_{nv}"] = {"us": 1e3 * ms, "GBs": bpe * nv / ms / 1e6, "Msteps_s": nv / ms / 1e3}
        del ve
print(json.dumps(out, indent=1))
