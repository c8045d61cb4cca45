"""Planner-level golden statistics for PredatorPrey-v0 and LevelBasedForaging-v3 from the
UNMODIFIED reference planner (/root/reference/posggym_baselines/planning, imported through
oracle.shims) running on the CPU oracle models.  Run in the build container:
    python tests/golden/make_planner_env_golden.py
Writes tests/golden/planner_envs.json (root statistics after N simulations from a first
observation; MCTS._rollout return distributions).  tests/test_pomcp_gpu.py compares the CUDA
planner with these."""
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
from posggym_baselines.planning.belief import HistoryPolicyState  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
ENVS = {
    "PredatorPrey-v0": dict(grid="10x10", num_predators=2, num_prey=3, cooperative=True, prey_strength=2, obs_dim=2),
    "LevelBasedForaging-v3": dict(num_agents=2, max_agent_level=3, size=10, max_food=8, sight=2, force_coop=False,
                                  static_layout=True, observation_mode="tuple"),
}


def make_config(seed):
    return MCTSConfig(discount=0.95, search_time_limit=0.1, c=1.414, truncated=False, action_selection="ucb",
                      state_belief_only=True, seed=seed)


def root_stats(env_id, env_seeds, n_sims=192, reps=100):
    out = []
    for scen, env_seed in enumerate(env_seeds):
        env = posggym.make(env_id, **ENVS[env_id])
        obs, _ = env.reset(seed=env_seed)
        values, visits = [], []
        for rep in range(reps):
            random.seed(7000 * scen + rep)
            planner = POMCP(env.model, "0", make_config(rep), search_policy=RandomSearchPolicy(env.model, "0"))
            env.model.seed(9000 + rep)
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
        values, visits = np.array(values), np.array(visits)
        out.append({
            "env_seed": env_seed, "n_sims": n_sims, "reps": reps, "ego_obs": [int(v) for v in obs["0"]],
            "value_mean": values.mean(0).tolist(), "value_std": values.std(0).tolist(),
            "visit_frac_mean": (visits / visits.sum(1, keepdims=True)).mean(0).tolist(),
        })
        print(env_id, "root stats", scen, np.round(out[-1]["value_mean"], 4), np.round(out[-1]["visit_frac_mean"], 3),
              flush=True)
    return out


def rollout_distribution(env_id, env_seed, fields_fn, n=4000):
    env = posggym.make(env_id, **ENVS[env_id])
    env.reset(seed=env_seed)
    state0 = env.state
    planner = POMCP(env.model, "0", make_config(0), search_policy=RandomSearchPolicy(env.model, "0"))
    env.model.seed(78)
    rets = []
    for _ in range(n):
        rets.append(planner._rollout(HistoryPolicyState(state0, None, None, 1), 0, planner.search_policy, {}))
    print(env_id, "rollouts", np.mean(rets), np.std(rets), flush=True)
    return {"env_seed": env_seed, "state": fields_fn(state0), "discount": 0.95, "depth_limit": planner.config.depth_limit,
            "step_limit": planner.step_limit, "returns_sorted": np.sort(np.array(rets)).round(12).tolist()}


if __name__ == "__main__":
    from tests.golden.make_golden import lbf_fields, pp_fields

    res = {"source": "unmodified reference planner on the CPU oracle (tests/golden/make_planner_env_golden.py)",
           "env_kwargs": ENVS,
           "PredatorPrey-v0": {"root_stats": root_stats("PredatorPrey-v0", (2, 5)),
                               "rollouts": rollout_distribution("PredatorPrey-v0", 2, lambda s: [int(v) for v in pp_fields(s)])},
           "LevelBasedForaging-v3": {"root_stats": root_stats("LevelBasedForaging-v3", (1, 4, 9)),
                                     "rollouts": rollout_distribution("LevelBasedForaging-v3", 1,
                                                                      lambda s: [int(v) for v in lbf_fields(s)])}}
    with open(os.path.join(HERE, "planner_envs.json"), "w") as f:
        json.dump(res, f)
    print("written")
