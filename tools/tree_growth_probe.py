"""Dev probe: root statistics at a large budget for small fixed wave sizes (is wave 1 the reference?)."""
import os, sys, json, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from posggym_baselines_b200.envs.env import Env, make
from posggym_baselines_b200.planning import POMCP, MCTSConfig, RandomSearchPolicy
G = os.path.join(ROOT, "tests", "golden")
drv = json.load(open(os.path.join(G, "planner_driving.json")))
def f2o(f):
    return (tuple(f[:15]), f[15], (f[16], f[17]), (f[18], f[19]), f[20], f[21])
model = make("Driving-v1").model
obs = f2o(next(r for r in drv["root_stats"] if r["env_seed"] == 3)["ego_obs"])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
for wave in [int(x) for x in sys.argv[2].split(',')]:
    for rep in range(4):
        cfg = MCTSConfig(discount=0.95, search_time_limit=0.1, c=1.414, truncated=False, action_selection="ucb", seed=rep)
        p = POMCP(model, "0", cfg, RandomSearchPolicy(model, "0"), max_sims=n, wave_size=wave, node_capacity=1 << 18, log_capacity=1 << 22)
        p.reset(); p.update(None, obs); t0 = time.time(); p.get_action(); dt = time.time() - t0
        st = p._last_root_stats
        print("wave", wave, "rep", rep, "max_depth", st["max_depth"], "visit share", np.round(st["done"] / st["done"].sum(), 3), "values",
              np.round(np.where(st["done"] > 0, st["total"] / np.maximum(st["done"], 1), 0), 4), "minmax", round(st["min_value"], 3), round(st["max_value"], 3), f"{dt:.2f} s", flush=True)
        p.close()
