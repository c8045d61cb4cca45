"""Dev probe for ncu: one chained launch of small search waves on a 65 536-visit tree."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from posggym_baselines_b200.envs.env import make
from posggym_baselines_b200.planning import POMCP, MCTSConfig, RandomSearchPolicy
G = os.path.join(ROOT, "tests", "golden")
drv = json.load(open(os.path.join(G, "planner_driving.json")))
f = next(r for r in drv["root_stats"] if r["env_seed"] == 3)["ego_obs"]
obs = (tuple(f[:15]), f[15], (f[16], f[17]), (f[18], f[19]), f[20], f[21])
model = make("Driving-v1").model
cfg = MCTSConfig(discount=0.95, search_time_limit=0.1, c=1.414, truncated=False, action_selection="ucb", seed=0)
p = POMCP(model, "0", cfg, RandomSearchPolicy(model, "0"), max_sims=int(sys.argv[1]) if len(sys.argv) > 1 else 32768,
          node_capacity=1 << 18, log_capacity=1 << 22)
p.reset(); p.update(None, obs); p.get_action()
print(p.search_info)
