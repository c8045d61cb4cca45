"""Root statistics of the UNMODIFIED reference POTMMCP (planning/potmmcp.py) with the reference's own
``OtherAgentMixturePolicy`` and ``POTMMCPMetaPolicy`` on the CPU oracle model, for two consecutive decisions:
the first (initial belief, root prior = expected action distribution of the meta-policy) and the second after
``update(action, obs)`` (subtree kept or node re-created, reinvigoration with typed policy states).  The policy
objects are the ones in tests/potmmcp_policies.py.  Writes tests/golden/potmmcp_driving.json."""
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
from posggym_baselines.planning import POTMMCP, MCTSConfig, OtherAgentMixturePolicy, POTMMCPMetaPolicy  # noqa: E402

from tests.golden.make_golden import obs_fields  # noqa: E402
from tests.potmmcp_policies import META_POLICY, ego_policies, other_agent_types  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def search(planner, n_sims):
    if not planner.root.children:
        for a in planner.action_space:
            planner.root.add_child(a)
    for _ in range(n_sims):                       # POTMMCP.get_action's loop with a simulation count for a clock
        hps = planner.root.belief.sample()
        pi = planner.search_policy.sample_policy(hps.policy_state)
        planner._simulate(hps, planner.root, 0, pi)
        planner.root.visits += 1
    return ([planner.root.children[a].value for a in planner.action_space],
            [planner.root.children[a].visits for a in planner.action_space],
            [planner.root.action_probs[a] for a in planner.action_space])


def summary(rows):
    values, visits, priors = (np.array(x, dtype=float) for x in zip(*rows))
    frac = visits / visits.sum(1, keepdims=True)
    return {"value_mean": values.mean(0).tolist(), "value_std": values.std(0).tolist(),
            "visit_frac_mean": frac.mean(0).tolist(), "visit_frac_std": frac.std(0).tolist(),
            "prior_mean": priors.mean(0).tolist(), "prior_std": priors.std(0).tolist()}


if __name__ == "__main__":
    out = []
    for truncated, n_sims, reps in ((False, 192, 64), (True, 512, 64)):
        env = posggym.make("Driving-v1", grid="14x14RoundAbout", num_agents=2, obs_dim=(3, 1, 1))
        obs0, _ = env.reset(seed=11)
        ego_action, other_action = 1, 1
        obs1 = env.step({"0": ego_action, "1": other_action})[0]
        first, second, types = [], [], []
        for rep in range(reps):
            random.seed(rep)
            env.model.seed(5000 + rep)
            cfg = MCTSConfig(discount=0.95, search_time_limit=0.1, c=1.25, truncated=truncated, action_selection="pucb",
                             pucb_exploration_fraction=0.5, state_belief_only=False, seed=rep)
            value = 0.05 if truncated else None
            ego = ego_policies(env.model)
            for pi in ego.values():
                pi.value = value
            planner = POTMMCP(env.model, "0", cfg,
                              {"1": OtherAgentMixturePolicy(env.model, "1", other_agent_types(env.model))},
                              POTMMCPMetaPolicy(env.model, "0", ego, META_POLICY))
            planner.reset()
            planner.update(None, obs0["0"])
            first.append(search(planner, n_sims))
            planner.update(ego_action, obs1["0"])
            ids = [p.policy_state["1"]["policy_id"] for p in planner.root.belief.particles]
            types.append(sum(i == "pusher" for i in ids) / len(ids))
            second.append(search(planner, n_sims))
        out.append({"truncated": truncated, "n_sims": n_sims, "reps": reps, "env_seed": 11, "ego_action": ego_action,
                    "obs0": obs_fields(obs0["0"]), "obs1": obs_fields(obs1["0"]),
                    "first": summary(first), "second": summary(second),
                    "pusher_share_mean": float(np.mean(types)), "pusher_share_std": float(np.std(types))})
        print(json.dumps(out[-1], indent=1), flush=True)
    json.dump({"source": "unmodified reference POTMMCP on the CPU oracle", "scenarios": out},
              open(os.path.join(HERE, "potmmcp_driving.json"), "w"))
