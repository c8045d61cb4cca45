"""Planner-level golden statistics from the UNMODIFIED reference planner
(/root/reference/posggym_baselines/planning, imported through oracle.shims) running on
the CPU oracle model.  Run in the build container:
    python tests/golden/make_planner_golden.py
Writes tests/golden/planner_driving.json.  The -m gpu tests compare the CUDA planner's
statistics with these (the reference tree does not exist on the GPU box)."""
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
from posggym_baselines.planning import POMCP, MCTSConfig, RandomSearchPolicy  # noqa: E402
from posggym_baselines.planning.belief import HistoryPolicyState  # noqa: E402

from tests.golden.make_golden import obs_fields, vs_fields  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
ENV_KW = dict(grid="14x14RoundAbout", num_agents=2, obs_dim=(3, 1, 1))


def make_config(seed, **kw):
    base = dict(discount=0.95, search_time_limit=0.1, c=1.414, truncated=False, action_selection="ucb",
                state_belief_only=True, seed=seed)
    base.update(kw)
    return MCTSConfig(**base)


def root_stats_scenarios(n_sims=192, reps=100):
    """Fixed first observation; N simulations of the get_action loop body (mcts.py:286-290);
    root action values / visit counts over `reps` independent planners."""
    out = []
    for scen, env_seed in enumerate((3, 11, 29)):
        env = posggym.make("Driving-v1", **ENV_KW)
        obs, _ = env.reset(seed=env_seed)
        state0 = env.state
        values, visits = [], []
        for rep in range(reps):
            random.seed(1000 * scen + rep)
            planner = POMCP(env.model, "0", make_config(rep), search_policy=RandomSearchPolicy(env.model, "0"))
            env.model.seed(5000 + rep)
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
            "env_seed": env_seed, "n_sims": n_sims, "reps": reps,
            "state": [vs_fields(v) for v in state0], "ego_obs": obs_fields(obs["0"]),
            "value_mean": values.mean(0).tolist(), "value_std": values.std(0).tolist(),
            "visit_frac_mean": (visits / visits.sum(1, keepdims=True)).mean(0).tolist(),
            "best_action_hist": np.bincount(values.argmax(1), minlength=5).tolist(),
        })
        print("root stats", scen, out[-1]["value_mean"], out[-1]["visit_frac_mean"], flush=True)
    return out


def rollout_distributions(n=3000):
    """Returns of MCTS._rollout (mcts.py:405-452) from a fixed state with the reference's
    own RNGs (gymnasium space sampling + model.rng)."""
    out = []
    for env_seed in (3, 11):
        env = posggym.make("Driving-v1", **ENV_KW)
        env.reset(seed=env_seed)
        state0 = env.state
        planner = POMCP(env.model, "0", make_config(0), search_policy=RandomSearchPolicy(env.model, "0"))
        env.model.seed(77)
        rets = []
        for _ in range(n):
            hps = HistoryPolicyState(state0, None, None, 1)
            rets.append(planner._rollout(hps, 0, planner.search_policy, {}))
        out.append({"env_seed": env_seed, "state": [vs_fields(v) for v in state0], "discount": 0.95,
                    "depth_limit": planner.config.depth_limit, "step_limit": planner.step_limit,
                    "returns_sorted": np.sort(np.array(rets)).round(12).tolist()})
        print("rollouts", env_seed, np.mean(rets), np.std(rets), flush=True)
    return out


def episodes(n_episodes=40):
    """README configuration (README.md:52-88): POMCP vs a uniform-random other agent,
    search_time_limit = 0.1 s; per-episode return / outcome / sims per decision."""
    env = posggym.make("Driving-v1", **ENV_KW)
    rets, lens, sims, outcomes = [], [], [], []
    t0 = time.time()
    for ep in range(n_episodes):
        planner = POMCP(env.model, "0", make_config(ep), search_policy=RandomSearchPolicy(env.model, "0"))
        other = pga.make("Random-v0", env.model, "1")
        obs, _ = env.reset(seed=ep)
        planner.reset()
        done, R, L = False, 0.0, 0
        while not done:
            a = planner.step(obs["0"])
            b = other.step(obs["1"])
            obs, rew, term, trunc, done, _ = env.step({"0": a, "1": b})
            R += rew["0"]
            L += 1
            if not math.isnan(planner.step_statistics["num_sims"]):
                sims.append(planner.step_statistics["num_sims"])
        vs = env.state[0]
        outcomes.append("WIN" if vs.dest_reached else ("LOSS" if vs.crashed else "NA"))
        rets.append(R)
        lens.append(L)
    print("episodes", np.mean(rets), np.mean(sims), time.time() - t0, flush=True)
    return {"n_episodes": n_episodes, "search_time_limit": 0.1, "returns": rets, "lens": lens,
            "outcomes": outcomes, "mean_sims_per_decision": float(np.mean(sims)),
            "host": f"{os.cpu_count()} cores, build container"}


if __name__ == "__main__":
    res = {"source": "unmodified reference planner on the CPU oracle (tests/golden/make_planner_golden.py)",
           "root_stats": root_stats_scenarios(), "rollouts": rollout_distributions(), "episodes": episodes()}
    with open(os.path.join(HERE, "planner_driving.json"), "w") as f:
        json.dump(res, f)
    print("written")
