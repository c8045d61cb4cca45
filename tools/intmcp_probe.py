"""Dev probe: I-NTMCP level 1 simulations per 0.2 s decision vs arena size."""
import math
import random
import sys

sys.path.insert(0, ".")
from posggym_baselines_b200.envs.driving import DrivingModel
from posggym_baselines_b200.envs.env import Env
from posggym_baselines_b200.planning import INTMCP, MCTSConfig

model = DrivingModel()
for cap_n, cap_l, mw in ((1 << 23, 1 << 25, 65536), (1 << 25, 1 << 27, 65536), (1 << 25, 1 << 27, 262144)):
    env = Env(model)
    cfg = MCTSConfig(discount=0.95, search_time_limit=0.2, c=math.sqrt(2), truncated=False, action_selection="ucb",
                     state_belief_only=False, seed=0)
    p = INTMCP.initialize(model, "0", cfg, nesting_level=1, search_policies=None, node_capacity=cap_n,
                          log_capacity=cap_l, max_wave=mw)
    obs, _ = env.reset(seed=200)
    p.reset()
    rng = random.Random(50)
    out = []
    for d in range(6):
        a = p.step(obs["0"])
        st = p.step_statistics
        out.append((int(st["num_sims"]), round(st["search_time"], 3), round(st["update_time"], 3)))
        obs, _, _, _, done, _ = env.step({"0": a, "1": rng.randrange(5)})
        if done:
            break
    p.close()
    print(cap_n, cap_l, mw, out)
