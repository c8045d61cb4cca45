"""Dev probe: where MCTS.update spends its time (times every C-ABI call with a device sync)."""
import collections
import random
import sys
import time

import torch

sys.path.insert(0, ".")
from posggym_baselines_b200 import _lib
from posggym_baselines_b200.envs.driving import DrivingModel
from posggym_baselines_b200.envs.env import Env
from posggym_baselines_b200.planning import POMCP, MCTSConfig, RandomSearchPolicy

lib = _lib.lib()
acc = collections.defaultdict(float)


class Timed:
    def __init__(self, name):
        self.name, self.fn = name, getattr(lib, name)

    def __call__(self, *a):
        torch.cuda.synchronize(); t0 = time.time()
        rc = self.fn(*a)
        torch.cuda.synchronize(); acc[self.name] += time.time() - t0
        return rc


class Proxy:
    def __getattr__(self, n):
        return Timed(n) if n.startswith("posg_") and n not in ("posg_last_error",) else getattr(lib, n)


_lib.lib = lambda: Proxy()
model = DrivingModel()
env = Env(model)
cfg = MCTSConfig(discount=0.95, search_time_limit=0.1, c=1.414, truncated=False, action_selection="ucb", seed=1)
p = POMCP(model, "0", cfg, RandomSearchPolicy(model, "0"))
obs, _ = env.reset(seed=1)
p.reset()
rng = random.Random(0)
for d in range(6):
    acc.clear()
    t0 = time.time()
    a = p.step(obs["0"])
    st = p.step_statistics
    print(d, f"wall {time.time()-t0:.3f} update {st['update_time']:.3f} reinv {st['reinvigoration_time']:.3f}",
          {k: round(v, 4) for k, v in acc.items() if k != "posg_search" and v > 0.0005}, "root particles", len(p._root_particles))
    obs, _, _, _, done, _ = env.step({"0": a, "1": rng.randrange(5)})
    if done:
        break
