import sys, random
sys.path.insert(0, ".")
from posggym_baselines_b200.envs.driving import DrivingModel
from posggym_baselines_b200.envs.env import Env
from posggym_baselines_b200.planning import POMCP, MCTSConfig, RandomSearchPolicy
model = DrivingModel(); env = Env(model)
cfg = MCTSConfig(discount=0.95, search_time_limit=0.1, c=1.414, truncated=False, action_selection="ucb", seed=1)
p = POMCP(model, "0", cfg, RandomSearchPolicy(model, "0"), max_sims=1 << 21, wave_size=65536, node_capacity=1 << 23, log_capacity=1 << 25)
obs, _ = env.reset(seed=1); p.reset()
a = p.step(obs["0"]); print(p.step_statistics["num_sims"], p.step_statistics["search_time"])
