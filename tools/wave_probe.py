"""Dev probe: POMCP simulations per 0.1 s decision vs the planner's maximum wave size."""
import random
import sys

sys.path.insert(0, ".")
from posggym_baselines_b200.envs.driving import DrivingModel
from posggym_baselines_b200.envs.env import Env
from posggym_baselines_b200.planning import POMCP, MCTSConfig, RandomSearchPolicy

model = DrivingModel()
for mw in (65536, 262144, 1048576):
    env = Env(model)
    cfg = MCTSConfig(discount=0.95, search_time_limit=0.1, c=1.414, truncated=False, action_selection="ucb", seed=1)
    p = POMCP(model, "0", cfg, RandomSearchPolicy(model, "0"), max_wave=mw)
    obs, _ = env.reset(seed=1)
    p.reset()
    rng = random.Random(0)
    out = []
    for d in range(6):
        a = p.step(obs["0"])
        out.append(int(p.step_statistics["num_sims"]))
        obs, _, _, _, done, _ = env.step({"0": a, "1": rng.randrange(5)})
        if done:
            break
    p.close()
    print(mw, out, sum(out[1:]) / max(1, len(out) - 1))
