"""The CUDA I-NTMCP planner (nesting level 1): API smoke on the reference's test scenario
(tests/planning/test_intmcp.py:34-73) and STATISTICAL parity with the unmodified reference
INTMCP at equal simulation counts per level (golden: tests/golden/make_intmcp_golden.py)."""
import json
import math
import os
import random

import numpy as np
import pytest

from tests.conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(GOLDEN, "intmcp_driving.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def model():
    from posggym_baselines_b200.envs.driving import DrivingModel

    return DrivingModel()


def make_planner(model, seed, epsilon=0.01, **opts):
    from posggym_baselines_b200.planning import INTMCP, MCTSConfig

    cfg = MCTSConfig(discount=0.95, search_time_limit=0.2, c=math.sqrt(2), truncated=False, action_selection="ucb",
                     pucb_exploration_fraction=0.25, epsilon=epsilon, seed=seed, state_belief_only=False)
    return INTMCP.initialize(model, "0", cfg, nesting_level=1, search_policies=None, **opts)


def fields_to_obs(f):
    return (tuple(f[:15]), f[15], (f[16], f[17]), (f[18], f[19]), f[20], f[21])


def test_reference_scenario_runs(model):
    from posggym_baselines_b200.envs.env import Env
    from posggym_baselines_b200.planning import PlanningStatTracker

    env = Env(model)
    planner = make_planner(model, seed=0, epsilon=0.92, node_capacity=1 << 21, log_capacity=1 << 23)   # depth_limit = 2
    assert planner.config.depth_limit == 2
    obs, _ = env.reset(seed=2)
    planner.reset()
    rng = random.Random(2)
    done, steps = False, 0
    while not done:
        a = planner.step(obs["0"])
        assert a in range(5)
        st = planner.step_statistics
        assert set(st) == set(PlanningStatTracker.STAT_KEYS)
        if not math.isnan(st["num_sims"]) and not planner._absorbing:
            assert st["num_sims"] > 1000
        obs, _, _, _, done, _ = env.step({"0": a, "1": rng.randrange(5)})
        steps += 1
    assert steps <= 50
    planner.close()


@pytest.mark.parametrize("wave", [1, 16])
def test_root_statistics_match_reference(model, gold, wave):
    """Root action values / visit shares of the level-1 tree after N level-0 and N level-1
    simulations from the same first observation: means within 4 SE + 0.02 / 0.05 of the
    reference planner's."""
    for scen in gold["root_stats"]:
        reps, n = 40, scen["n_sims_per_level"]
        vals, vis = [], []
        for rep in range(reps):
            p = make_planner(model, seed=rep, epsilon=scen["epsilon"], max_sims_per_level=n, wave_size=wave,
                             node_capacity=1 << 16, log_capacity=1 << 19)
            p.reset()
            p.update(None, fields_to_obs(scen["ego_obs"]))
            p.get_action()
            st = p._last_root_stats
            vals.append(np.where(st["done"] > 0, st["total"] / np.maximum(st["done"], 1), 0.0))
            vis.append(st["done"] / max(1.0, st["done"].sum()))
            p.close()
        vals, vis = np.array(vals), np.array(vis)
        ref_mean, ref_std = np.array(scen["value_mean"]), np.array(scen["value_std"])
        se = np.sqrt(ref_std**2 / scen["reps"] + vals.std(0) ** 2 / reps)
        assert (np.abs(vals.mean(0) - ref_mean) < 4 * se + 0.02).all(), (vals.mean(0), ref_mean, se)
        assert (np.abs(vis.mean(0) - np.array(scen["visit_frac_mean"])) < 0.05).all(), (vis.mean(0), scen["visit_frac_mean"])


def test_episode_outcomes_match_reference(model, gold):
    from posggym_baselines_b200.envs.env import Env

    ref = gold["episodes"]
    n_sims = max(8, int(round(ref["mean_sims_per_decision"] / 2)))
    env = Env(model)
    wins, n = 0, 40
    for ep in range(n):
        planner = make_planner(model, seed=ep, max_sims_per_level=n_sims, wave_size=1, node_capacity=1 << 16,
                               log_capacity=1 << 19)
        rng = random.Random(20_000 + ep)
        obs, _ = env.reset(seed=ep)
        planner.reset()
        done = False
        while not done:
            a = planner.step(obs["0"])
            obs, _, _, _, done, _ = env.step({"0": a, "1": rng.randrange(5)})
        wins += env.state[0].dest_reached
        planner.close()
    p_ref = ref["outcomes"].count("WIN") / ref["n_episodes"]
    p = wins / n
    se = math.sqrt(p_ref * (1 - p_ref) / ref["n_episodes"] + p * (1 - p) / n)
    assert abs(p - p_ref) < 3 * se + 0.07, (p, p_ref)


@pytest.mark.parametrize("name", ["level2_random", "level1_plugins"])
def test_general_hierarchies_match_reference(model, name):
    """Beyond BASELINE's configuration -- nesting level 2, and plugin search policies at nesting level 1 --
    ``INTMCP.initialize`` builds the hierarchy on the plugin path (planning/host_intmcp.py: host trees, initial
    beliefs / reinvigoration rounds / model steps on the device).  Root values and visit shares of two consecutive
    decisions and the number of nested histories match a golden of the UNMODIFIED reference INTMCP
    (tests/golden/make_intmcp_deep_golden.py) within 4 standard errors plus a small margin."""
    from posggym_baselines_b200.planning import INTMCP, MCTSConfig, SearchPolicyWrapper
    from posggym_baselines_b200.planning.host_intmcp import HostINTMCP
    from tests.potmmcp_policies import plugin_policies_for

    with open(os.path.join(GOLDEN, "intmcp_deep_driving.json")) as f:
        scen = next(s for s in json.load(f)["scenarios"] if s["name"] == name)
    obs0, obs1 = fields_to_obs(scen["obs0"]), fields_to_obs(scen["obs1"])
    level, reps = scen["nesting_level"], 6
    rows, nested = {"first": [], "second": []}, []
    for rep in range(reps):
        random.seed(4000 + rep)
        model.seed(9700 + rep)
        cfg = MCTSConfig(discount=0.95, search_time_limit=0.2, c=math.sqrt(2), truncated=False, action_selection="ucb",
                         state_belief_only=False, seed=rep)
        sp = plugin_policies_for(model, level, SearchPolicyWrapper) if scen["plugins"] else None
        planner = INTMCP.initialize(model, "0", cfg, level, sp, max_sims_per_level=scen["n_sims_per_level"])
        assert isinstance(planner, HostINTMCP)
        planner.reset()
        planner.update(None, obs0)
        for which in ("first", "second"):
            assert planner.get_action() in range(5)
            assert planner.step_statistics["num_sims"] == scen["n_sims_per_level"] * (level + 1)
            root = planner.traverse(planner.history)
            vis = np.array([root.children[a].visits for a in range(5)], dtype=float)
            rows[which].append(([root.children[a].value for a in range(5)], vis / vis.sum()))
            if which == "first":
                planner.update(scen["ego_action"], obs1)
                assert root.belief.size() == 22 and planner.traverse(planner.history).belief.size() >= 22
                nested.append(len(planner.get_nested_history_dist({planner.history: 1.0})["1"]))
        planner.close()
    for which in ("first", "second"):
        vals, frac = (np.array(x) for x in zip(*rows[which]))
        for got, key, margin in ((vals, "value", 0.015), (frac, "visit_frac", 0.04)):
            mean, std = np.array(scen[which][key + "_mean"]), np.array(scen[which][key + "_std"])
            se = np.sqrt(std**2 / scen["reps"] + got.std(0) ** 2 / reps)
            assert (np.abs(got.mean(0) - mean) < 4 * se + margin).all(), (which, key, got.mean(0), mean, se)
    se = math.sqrt(scen["nested_histories_std"] ** 2 / scen["reps"] + np.std(nested) ** 2 / reps)
    assert abs(np.mean(nested) - scen["nested_histories_mean"]) < 4 * se + 0.75
