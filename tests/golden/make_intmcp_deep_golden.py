"""Root statistics of the UNMODIFIED reference INTMCP beyond BASELINE's configuration, on the CPU oracle:
(a) nesting level 2 with random search policies, (b) nesting level 1 with plugin search policies
(``SearchPolicyWrapper`` around the policy objects of tests/potmmcp_policies.py) -- two consecutive decisions each
(the second after ``update(action, obs)``: nested history distributions, reinvigoration through the trees one level
down).  Writes tests/golden/intmcp_deep_driving.json."""
from __future__ import annotations

import json
import math
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import shims  # noqa: E402

assert shims.install(), "needs /root/reference"
import posggym  # noqa: E402
from posggym_baselines.planning import INTMCP, MCTSConfig, SearchPolicyWrapper  # noqa: E402

from tests.golden.make_golden import obs_fields  # noqa: E402
from tests.potmmcp_policies import plugin_policies_for  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def search(planner, n_sims):
    for level in range(planner.nesting_level + 1):      # get_action's loop (intmcp.py:385-392) with fixed counts
        for _ in range(n_sims):
            planner._nested_sim(planner.history, level, True)
    root = planner.traverse(planner.history)
    vis = np.array([root.children[a].visits for a in range(5)], dtype=float)
    return [root.children[a].value for a in range(5)], (vis / vis.sum()).tolist(), root.belief.size()


def summary(rows):
    values, frac, size = zip(*rows)
    values, frac = np.array(values), np.array(frac)
    return {"value_mean": values.mean(0).tolist(), "value_std": values.std(0).tolist(),
            "visit_frac_mean": frac.mean(0).tolist(), "visit_frac_std": frac.std(0).tolist(),
            "belief_size_mean": float(np.mean(size))}


if __name__ == "__main__":
    out = []
    for name, level, plugins, n_sims, reps in (("level2_random", 2, False, 64, 48), ("level1_plugins", 1, True, 96, 48)):
        env = posggym.make("Driving-v1", grid="14x14RoundAbout", num_agents=2, obs_dim=(3, 1, 1))
        obs0, _ = env.reset(seed=11)
        ego_action = 1
        obs1 = env.step({"0": ego_action, "1": 1})[0]
        first, second, nested = [], [], []
        for rep in range(reps):
            random.seed(rep)
            env.model.seed(9000 + rep)
            cfg = MCTSConfig(discount=0.95, search_time_limit=0.2, c=math.sqrt(2), truncated=False,
                             action_selection="ucb", state_belief_only=False, seed=rep)
            sp = plugin_policies_for(env.model, level, SearchPolicyWrapper) if plugins else None
            planner = INTMCP.initialize(env.model, "0", cfg, nesting_level=level, search_policies=sp)
            planner.reset()
            planner.update(None, obs0["0"])
            first.append(search(planner, n_sims))
            planner.update(ego_action, obs1["0"])
            nested.append(len(planner.get_nested_history_dist({planner.history: 1.0})["1"]))
            second.append(search(planner, n_sims))
        out.append({"name": name, "nesting_level": level, "plugins": plugins, "n_sims_per_level": n_sims, "reps": reps,
                    "env_seed": 11, "ego_action": ego_action, "obs0": obs_fields(obs0["0"]), "obs1": obs_fields(obs1["0"]),
                    "first": summary(first), "second": summary(second),
                    "nested_histories_mean": float(np.mean(nested)), "nested_histories_std": float(np.std(nested))})
        print(name, json.dumps({k: out[-1][k] for k in ("first", "second", "nested_histories_mean")}), flush=True)
    json.dump({"source": "unmodified reference INTMCP on the CPU oracle", "scenarios": out},
              open(os.path.join(HERE, "intmcp_deep_driving.json"), "w"))
