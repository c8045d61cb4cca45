"""Large-budget golden statistics from the UNMODIFIED reference planner (imported through oracle.shims,
running on the CPU oracle models): root action values / visit shares after 65 536 simulations from a
fixed first observation -- the budget regime of the CUDA planner's throughput mode (wave sizes doubling
from 32 to tens of thousands), which the 192-simulation goldens of make_planner_golden.py do not reach.
Run in the build container (a few minutes on 8 cores):
    python tests/golden/make_planner_large_golden.py
Writes tests/golden/planner_large.json; tests/test_pomcp_gpu.py compares the CUDA planner with it."""
from __future__ import annotations

import json
import os
import random
import sys
import time
from multiprocessing import get_context

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

HERE = os.path.dirname(os.path.abspath(__file__))
N_SIMS, REPS = 65536, 4
SCENARIOS = [("Driving-v1", 3), ("Driving-v1", 11), ("PredatorPrey-v0", 2), ("LevelBasedForaging-v3", 1)]
ENVS = {
    "Driving-v1": dict(grid="14x14RoundAbout", num_agents=2, obs_dim=(3, 1, 1)),
    "PredatorPrey-v0": dict(grid="10x10", num_predators=2, num_prey=3, cooperative=True, prey_strength=2, obs_dim=2),
    "LevelBasedForaging-v3": dict(num_agents=2, max_agent_level=3, size=10, max_food=8, sight=2, force_coop=False,
                                  static_layout=True, observation_mode="tuple"),
}


def run(args):
    env_id, env_seed, rep = args
    from oracle import shims

    assert shims.install(), "needs /root/reference"
    import posggym
    from posggym_baselines.planning import POMCP, MCTSConfig, RandomSearchPolicy

    env = posggym.make(env_id, **ENVS[env_id])
    obs, _ = env.reset(seed=env_seed)
    random.seed(31000 + 100 * env_seed + rep)
    cfg = MCTSConfig(discount=0.95, search_time_limit=0.1, c=1.414, truncated=False, action_selection="ucb",
                     state_belief_only=True, seed=rep)
    planner = POMCP(env.model, "0", cfg, search_policy=RandomSearchPolicy(env.model, "0"))
    env.model.seed(41000 + rep)
    planner.reset()
    planner.update(None, obs["0"])
    for a in planner.action_space:
        planner.root.add_child(a)
    t0 = time.time()
    for _ in range(N_SIMS):
        hps = planner.root.belief.sample()
        planner._simulate(hps, planner.root, 0, planner.search_policy)
        planner.root.visits += 1
    values = [planner.root.children[a].value for a in planner.action_space]
    visits = [planner.root.children[a].visits for a in planner.action_space]
    return env_id, env_seed, rep, values, visits, time.time() - t0


if __name__ == "__main__":
    jobs = [(e, s, r) for e, s in SCENARIOS for r in range(REPS)]
    with get_context("fork").Pool(min(len(jobs), os.cpu_count() or 1)) as pool:
        res = pool.map(run, jobs, chunksize=1)
    out = {"source": "unmodified reference planner on the CPU oracle (tests/golden/make_planner_large_golden.py)",
           "n_sims": N_SIMS, "reps": REPS, "env_kwargs": ENVS, "scenarios": []}
    for env_id, env_seed in SCENARIOS:
        rows = [r for r in res if r[0] == env_id and r[1] == env_seed]
        values = np.array([r[3] for r in rows])
        visits = np.array([r[4] for r in rows], dtype=np.float64)
        out["scenarios"].append({
            "env_id": env_id, "env_seed": env_seed,
            "values": values.tolist(), "value_mean": values.mean(0).tolist(), "value_std": values.std(0).tolist(),
            "visit_frac": (visits / visits.sum(1, keepdims=True)).tolist(),
            "visit_frac_mean": (visits / visits.sum(1, keepdims=True)).mean(0).tolist(),
            "seconds_per_run": float(np.mean([r[5] for r in rows]))})
        print(env_id, env_seed, np.round(values.mean(0), 4), np.round(values.std(0), 4),
              np.round(out["scenarios"][-1]["visit_frac_mean"], 3), flush=True)
    with open(os.path.join(HERE, "planner_large.json"), "w") as f:
        json.dump(out, f)
    print("written")
