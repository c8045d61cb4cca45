"""I-NTMCP golden statistics from the UNMODIFIED reference planner
(/root/reference/posggym_baselines/planning/intmcp.py, imported through oracle.shims) on the
CPU oracle model.  Run in the build container: python tests/golden/make_intmcp_golden.py
Writes tests/golden/intmcp_driving.json."""
from __future__ import annotations

import json
import math
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import shims  # noqa: E402

assert shims.install(), "needs /root/reference"
import posggym  # noqa: E402
import posggym.agents as pga  # noqa: E402
from posggym_baselines.planning import INTMCP, MCTSConfig  # noqa: E402

from tests.golden.make_golden import obs_fields, vs_fields  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
ENV_KW = dict(grid="14x14RoundAbout", num_agents=2, obs_dim=(3, 1, 1))


def make_config(seed, **kw):
    base = dict(discount=0.95, search_time_limit=0.2, c=math.sqrt(2), truncated=False, action_selection="ucb",
                state_belief_only=False, seed=seed)
    base.update(kw)
    return MCTSConfig(**base)


def root_stats(n_sims=128, reps=80):
    out = []
    for scen, (env_seed, eps) in enumerate(((3, 0.01), (11, 0.01), (29, 0.92))):
        env = posggym.make("Driving-v1", **ENV_KW)
        obs, _ = env.reset(seed=env_seed)
        values, visits = [], []
        for rep in range(reps):
            random.seed(1000 * scen + rep)
            env.model.seed(7000 + rep)
            planner = INTMCP.initialize(env.model, "0", make_config(rep, epsilon=eps), nesting_level=1, search_policies=None)
            planner.reset()
            planner.update(None, obs["0"])
            for level in range(2):          # get_action's loop (intmcp.py:385-392) with fixed counts
                for _ in range(n_sims):
                    planner._nested_sim(planner.history, level, True)
            root = planner.traverse(planner.history)
            values.append([root.children[a].value for a in range(5)])
            visits.append([root.children[a].visits for a in range(5)])
        values, visits = np.array(values), np.array(visits, dtype=float)
        out.append({"env_seed": env_seed, "epsilon": eps, "n_sims_per_level": n_sims, "reps": reps,
                    "state": [vs_fields(v) for v in env.state], "ego_obs": obs_fields(obs["0"]),
                    "value_mean": values.mean(0).tolist(), "value_std": values.std(0).tolist(),
                    "visit_frac_mean": (visits / visits.sum(1, keepdims=True)).mean(0).tolist()})
        print("intmcp root", scen, out[-1]["value_mean"], out[-1]["visit_frac_mean"], flush=True)
    return out


def episodes(n_episodes=30):
    """tests/planning/test_intmcp.py:34-73 scenario (depth_limit 2, 0.2 s) vs a random other agent."""
    env = posggym.make("Driving-v1", **ENV_KW)
    rets, outcomes, sims = [], [], []
    t0 = time.time()
    for ep in range(n_episodes):
        planner = INTMCP.initialize(env.model, "0", make_config(ep, epsilon=0.01), nesting_level=1, search_policies=None)
        other = pga.make("Random-v0", env.model, "1")
        obs, _ = env.reset(seed=ep)
        planner.reset()
        done, R = False, 0.0
        while not done:
            a = planner.step(obs["0"])
            b = other.step(obs["1"])
            obs, rew, term, trunc, done, _ = env.step({"0": a, "1": b})
            R += rew["0"]
            if not math.isnan(planner.step_statistics["num_sims"]):
                sims.append(planner.step_statistics["num_sims"])
        vs = env.state[0]
        outcomes.append("WIN" if vs.dest_reached else ("LOSS" if vs.crashed else "NA"))
        rets.append(R)
    print("intmcp episodes", np.mean(rets), np.mean(sims), time.time() - t0, flush=True)
    return {"n_episodes": n_episodes, "search_time_limit": 0.2, "returns": rets, "outcomes": outcomes,
            "mean_sims_per_decision": float(np.mean(sims))}


if __name__ == "__main__":
    res = {"source": "unmodified reference INTMCP (nesting_level=1) on the CPU oracle",
           "root_stats": root_stats(), "episodes": episodes()}
    with open(os.path.join(HERE, "intmcp_driving.json"), "w") as f:
        json.dump(res, f)
    print("written")
