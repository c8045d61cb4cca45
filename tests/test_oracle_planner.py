"""The restated CPU planner (oracle/pomcp.py, the CPU baseline on the GPU box) against
golden statistics of the unmodified reference planner.  CPU only."""
import json
import os
import random

import numpy as np

from tests.conftest import GOLDEN


def test_oracle_pomcp_matches_reference_root_statistics(oracle_env):
    import posggym
    from oracle.pomcp import OraclePOMCP

    gold = json.load(open(os.path.join(GOLDEN, "planner_driving.json")))
    scen = gold["root_stats"][0]
    env = posggym.make("Driving-v1", grid="14x14RoundAbout", num_agents=2, obs_dim=(3, 1, 1))
    obs, _ = env.reset(seed=scen["env_seed"])
    vals, vis, reps = [], [], 30
    for rep in range(reps):
        random.seed(rep)
        env.model.seed(900 + rep)
        p = OraclePOMCP(env.model, "0", 0.95, 0.1, 1.414, seed=rep)
        p.reset()
        p.update(None, obs["0"])
        p.get_action(max_sims=scen["n_sims"])
        vals.append([p.root.children[a].value for a in p.actions])
        vis.append([p.root.children[a].visits for a in p.actions])
    vals, vis = np.array(vals), np.array(vis)
    se = np.sqrt(np.array(scen["value_std"]) ** 2 / scen["reps"] + vals.std(0) ** 2 / reps)
    assert (np.abs(vals.mean(0) - np.array(scen["value_mean"])) < 4 * se + 0.02).all()
    assert (np.abs((vis / vis.sum(1, keepdims=True)).mean(0) - np.array(scen["visit_frac_mean"])) < 0.05).all()


def test_config_derivations_match_reference():
    """MCTSConfig derived fields (planning/config.py:47-55) for the configs SURVEY.md 8 lists."""
    from posggym_baselines_b200.planning.config import MCTSConfig

    c = MCTSConfig(discount=0.95, search_time_limit=0.1, c=1.414, truncated=False, action_selection="UCB")
    assert (c.num_particles, c.extra_particles, c.depth_limit, c.action_selection) == (10, 1, 90, "ucb")
    c = MCTSConfig(discount=0.95, search_time_limit=0.1, c=1.414, truncated=False, epsilon=0.92)
    assert c.depth_limit == 2 and c.action_selection == "pucb"
    for t, n in ((0.1, 10), (1, 100), (5, 500), (10, 1000), (20, 2000)):
        c = MCTSConfig(discount=0.99, search_time_limit=t, c=1.25, truncated=True)
        assert c.num_particles == n and c.depth_limit == 459
    assert c.extra_particles == 125
    c2 = c.replace(state_belief_only=True)
    assert c2.state_belief_only and not c.state_belief_only and c2.num_particles == c.num_particles
