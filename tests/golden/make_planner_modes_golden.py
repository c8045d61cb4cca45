"""Root statistics of the UNMODIFIED reference POMCP for the other action-selection modes
(pucb -> max-visit final selection, uniform -> min-visit search; planning/mcts.py:66-74,
494-565) on the CPU oracle.  Writes tests/golden/planner_modes_driving.json."""
from __future__ import annotations

import json
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import shims  # noqa: E402

assert shims.install(), "needs /root/reference"
import posggym  # noqa: E402
from posggym_baselines.planning import POMCP, MCTSConfig, RandomSearchPolicy  # noqa: E402

from tests.golden.make_golden import obs_fields  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

if __name__ == "__main__":
    out = []
    for mode, c in (("pucb", 1.25), ("uniform", 1.414)):
        env = posggym.make("Driving-v1", grid="14x14RoundAbout", num_agents=2, obs_dim=(3, 1, 1))
        obs, _ = env.reset(seed=11)
        values, visits = [], []
        n_sims, reps = 192, 80
        for rep in range(reps):
            random.seed(rep)
            env.model.seed(3000 + rep)
            cfg = MCTSConfig(discount=0.95, search_time_limit=0.1, c=c, truncated=False, action_selection=mode,
                             pucb_exploration_fraction=0.5, state_belief_only=True, seed=rep)
            planner = POMCP(env.model, "0", cfg, search_policy=RandomSearchPolicy(env.model, "0"))
            planner.reset()
            planner.update(None, obs["0"])
            for a in planner.action_space:
                planner.root.add_child(a)
            for _ in range(n_sims):
                hps = planner.root.belief.sample()
                planner._simulate(hps, planner.root, 0, planner.search_policy)
                planner.root.visits += 1
            values.append([planner.root.children[a].value for a in planner.action_space])
            visits.append([planner.root.children[a].visits for a in planner.action_space])
        values, visits = np.array(values), np.array(visits, dtype=float)
        out.append({"mode": mode, "c": c, "env_seed": 11, "n_sims": n_sims, "reps": reps, "ego_obs": obs_fields(obs["0"]),
                    "value_mean": values.mean(0).tolist(), "value_std": values.std(0).tolist(),
                    "visit_frac_mean": (visits / visits.sum(1, keepdims=True)).mean(0).tolist(),
                    "visit_frac_std": (visits / visits.sum(1, keepdims=True)).std(0).tolist()})
        print(mode, out[-1]["value_mean"], out[-1]["visit_frac_mean"], flush=True)
    json.dump({"source": "unmodified reference POMCP on the CPU oracle", "modes": out},
              open(os.path.join(HERE, "planner_modes_driving.json"), "w"))
