"""Dev probe: POMCP sims / update time per decision vs arena size."""
import random
import sys

sys.path.insert(0, ".")
from posggym_baselines_b200.envs.driving import DrivingModel
from posggym_baselines_b200.envs.env import Env
from posggym_baselines_b200.planning import POMCP, MCTSConfig, RandomSearchPolicy

model = DrivingModel()
for cap_n, cap_l in ((1 << 25, 1 << 26), (1 << 26, 1 << 28), (1 << 27, 1 << 29)):
    env = Env(model)
    cfg = MCTSConfig(discount=0.95, search_time_limit=0.1, c=1.414, truncated=False, action_selection="ucb", seed=1)
    p = POMCP(model, "0", cfg, RandomSearchPolicy(model, "0"), node_capacity=cap_n, log_capacity=cap_l)
    obs, _ = env.reset(seed=1)
    p.reset()
    rng = random.Random(0)
    out = []
    for d in range(6):
        a = p.step(obs["0"])
        st = p.step_statistics
        out.append((round(st["num_sims"] / 1e6, 1), round(st["update_time"], 3), int(p._last_root_stats["node_count"]) >> 20, int(st["search_depth"])))
        obs, _, _, _, done, _ = env.step({"0": a, "1": rng.randrange(5)})
        if done:
            break
    p.close()
    print(cap_n >> 20, cap_l >> 20, out)
