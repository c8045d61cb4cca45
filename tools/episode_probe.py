"""Dev probe: per-decision planner statistics over a few decisions of a time-limited episode."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from posggym_baselines_b200.envs.env import Env, make
from posggym_baselines_b200.planning import POMCP, MCTSConfig, RandomSearchPolicy
import random
model = make("Driving-v1").model
env = Env(model)
limit = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
nd = int(sys.argv[2]) if len(sys.argv) > 2 else 6
cfg = MCTSConfig(discount=0.95, search_time_limit=limit, c=1.414, truncated=False, action_selection="ucb", seed=0)
p = POMCP(model, "0", cfg, RandomSearchPolicy(model, "0"))
obs, _ = env.reset(seed=100); p.reset()
rng = random.Random(0)
for d in range(nd):
    a = p.step(obs["0"])
    st = p._last_root_stats
    print(d, "action", a, {k: (round(v, 4) if isinstance(v, float) else v) for k, v in p.search_info.items()},
          "max_depth", st["max_depth"], "root visits", st["node_visits"], "belief", p.root.belief.size(),
          "update ms", round(1e3 * p.step_statistics["update_time"], 2), "kept nodes", p._nodes_bound, flush=True)
    obs, _, _, _, done, _ = env.step({"0": a, "1": rng.randrange(5)})
    if done: break
p.close()
