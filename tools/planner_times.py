"""Dev probe: per-decision statistics of the CUDA POMCP planner (README config)."""
import random
import sys

sys.path.insert(0, ".")
from posggym_baselines_b200.envs.driving import DrivingModel
from posggym_baselines_b200.envs.env import Env
from posggym_baselines_b200.planning import POMCP, MCTSConfig, RandomSearchPolicy

model = DrivingModel()
env = Env(model)
cfg = MCTSConfig(discount=0.95, search_time_limit=0.1, c=1.414, truncated=False, action_selection="ucb", seed=1)
p = POMCP(model, "0", cfg, RandomSearchPolicy(model, "0"))
obs, _ = env.reset(seed=1)
p.reset()
rng = random.Random(0)
import time
for d in range(8):
    t0 = time.time()
    a = p.step(obs["0"])
    dt = time.time() - t0
    st = p.step_statistics
    print(d, f"wall {dt:.3f}", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in st.items() if "time" in k or k in ("num_sims", "search_depth")},
          "compactions", p.tree_compactions)
    obs, _, _, _, done, _ = env.step({"0": a, "1": rng.randrange(5)})
    if done:
        break
