"""Dev probe: POMCP simulations per 0.1 s decision on Driving-v1 (README config), a few decisions."""
import random
import sys

sys.path.insert(0, ".")
from posggym_baselines_b200.envs.driving import DrivingModel
from posggym_baselines_b200.envs.env import Env
from posggym_baselines_b200.planning import POMCP, MCTSConfig, RandomSearchPolicy

model = DrivingModel()
env = Env(model)
cfg = MCTSConfig(discount=0.95, search_time_limit=0.1, c=1.414, truncated=False, action_selection="ucb", seed=1)
if len(sys.argv) > 1 and sys.argv[1] == "fixed":
    p = POMCP(model, "0", cfg, RandomSearchPolicy(model, "0"), max_sims=1 << 21, wave_size=65536,
              node_capacity=1 << 23, log_capacity=1 << 25)
else:
    p = POMCP(model, "0", cfg, RandomSearchPolicy(model, "0"))
obs, _ = env.reset(seed=1)
p.reset()
rng = random.Random(0)
out = []
for d in range(7):
    a = p.step(obs["0"])
    out.append((int(p.step_statistics["num_sims"]), round(p.step_statistics["search_time"], 4)))
    obs, _, _, _, done, _ = env.step({"0": a, "1": rng.randrange(5)})
    if done:
        break
print(out)
