"""HOST logic of the plugin-path planners (planning/host_search.py, planning/potmmcp.py, planning/host_intmcp.py) on
CPU: the engines run unchanged against tests/host_double.py -- a stand-in for the device facade's batch interface
answered by the CPU oracle model -- and their root statistics are compared with goldens of the UNMODIFIED reference
planners (tests/golden/make_potmmcp_golden.py, make_intmcp_deep_golden.py).  The same comparisons run with the real
device facade in tests/test_plugins_gpu.py / tests/test_intmcp_gpu.py."""
import json
import math
import os
import random

import numpy as np
import pytest

from tests.conftest import GOLDEN


def _driving_obs(f):
    return (tuple(f[:15]), f[15], (f[16], f[17]), (f[18], f[19]), f[20], f[21])


def _close(got, mean, std, reps_ref, margin):
    got, mean, std = np.array(got), np.array(mean), np.array(std)
    se = np.sqrt(std**2 / reps_ref + got.std(0) ** 2 / len(got))
    return (np.abs(got.mean(0) - mean) < 4 * se + margin).all(), (got.mean(0), mean, se)


@pytest.fixture()
def double_model(oracle_env):
    from tests.host_double import DoubleDrivingModel

    return DoubleDrivingModel()


class _PlannerShell:
    """What a host engine reads of its ``MCTS`` facade (planning/mcts.py), without the device search."""

    def __init__(self, model, agent_id, config, other_agent_policies, search_policy, max_sims):
        self.model, self.agent_id, self.config = model, agent_id, config
        self.other_agent_policies, self.search_policy, self.max_sims = other_agent_policies, search_policy, max_sims
        self.action_space = list(range(model.action_spaces[agent_id].n))
        self.step_limit = model.spec.max_episode_steps
        self.step_statistics = {"search_time": 0.0, "update_time": 0.0, "reinvigoration_time": 0.0, "evaluation_time": 0.0,
                                "policy_calls": 0, "inference_time": 0.0, "search_depth": 0, "num_sims": 0, "mem_usage": 0,
                                "min_value": 0.0, "max_value": 0.0}

    def _device_initial_belief(self, init_obs, need):
        return self.model.initial_belief(self.agent_id, init_obs, need)


@pytest.mark.parametrize("truncated", [False, True])
def test_potmmcp_engine_matches_reference_golden(double_model, truncated):
    from posggym_baselines_b200.planning import MCTSConfig, OtherAgentMixturePolicy, POTMMCPMetaPolicy
    from posggym_baselines_b200.planning.potmmcp import _PotmmcpEngine
    from tests.potmmcp_policies import META_POLICY, ego_policies, other_agent_types

    scen = next(s for s in json.load(open(os.path.join(GOLDEN, "potmmcp_driving.json")))["scenarios"]
                if s["truncated"] == truncated)
    obs0, obs1 = _driving_obs(scen["obs0"]), _driving_obs(scen["obs1"])
    model, reps = double_model, 8
    rows, shares = {"first": [], "second": []}, []
    for rep in range(reps):
        random.seed(2000 + rep)
        model.seed(8000 + rep)
        cfg = MCTSConfig(discount=0.95, search_time_limit=0.1, c=1.25, truncated=truncated, action_selection="pucb",
                         pucb_exploration_fraction=0.5, state_belief_only=False, seed=rep)
        ego = ego_policies(model)
        for pi in ego.values():
            pi.value = 0.05 if truncated else None
        shell = _PlannerShell(model, "0", cfg, {"1": OtherAgentMixturePolicy(model, "1", other_agent_types(model))},
                              POTMMCPMetaPolicy(model, "0", ego, META_POLICY), scen["n_sims"])
        eng = _PotmmcpEngine(shell)
        eng.update(None, obs0)
        for which in ("first", "second"):
            assert eng.get_action() in range(5) and shell.step_statistics["num_sims"] == scen["n_sims"]
            vis = np.array([eng.root.children[a].visits for a in range(5)], dtype=float)
            rows[which].append(([eng.root.children[a].value for a in range(5)], vis / vis.sum(),
                                [eng.root.action_probs[a] for a in range(5)]))
            if which == "first":
                eng.update(scen["ego_action"], obs1)
                parts = eng.root.belief.particles
                assert len(parts) >= cfg.num_particles + cfg.extra_particles and all(p.t == 2 for p in parts)
                shares.append(np.mean([p.policy_state["1"]["policy_id"] == "pusher" for p in parts]))
    for which in ("first", "second"):
        vals, frac, prior = zip(*rows[which])
        for got, key, margin in ((vals, "value", 0.01), (frac, "visit_frac", 0.03), (prior, "prior", 0.01)):
            ok, info = _close(got, scen[which][key + "_mean"], scen[which][key + "_std"], scen["reps"], margin)
            assert ok, (which, key, info)
        # the search keeps folding the sampled policies' distributions into the root prior: it must MOVE between
        # repetitions as it does in the reference (a prior that is never updated has zero spread)
        assert np.array(prior)[:, 1].std() > 0.4 * scen[which]["prior_std"][1]
    se = math.sqrt(scen["pusher_share_std"] ** 2 / scen["reps"] + np.std(shares) ** 2 / reps)
    assert abs(np.mean(shares) - scen["pusher_share_mean"]) < 4 * se + 0.05


@pytest.mark.parametrize("name", ["level2_random", "level1_plugins"])
def test_general_intmcp_matches_reference_golden(double_model, name):
    """Nesting level 2 (three trees deep) and plugin search policies: root values, visit shares and belief sizes of
    two consecutive decisions, and the number of nested histories the update hands to the tree below."""
    from posggym_baselines_b200.planning import MCTSConfig, SearchPolicyWrapper
    from posggym_baselines_b200.planning.host_intmcp import HostINTMCP
    from tests.potmmcp_policies import plugin_policies_for

    class OnDouble(HostINTMCP):
        def _device_initial_belief(self, init_obs, need):
            return self.model.initial_belief(self.agent_id, init_obs, need)

    scen = next(s for s in json.load(open(os.path.join(GOLDEN, "intmcp_deep_driving.json")))["scenarios"] if s["name"] == name)
    obs0, obs1 = _driving_obs(scen["obs0"]), _driving_obs(scen["obs1"])
    model, reps, level = double_model, 8, scen["nesting_level"]
    rows, nested = {"first": [], "second": []}, []
    for rep in range(reps):
        random.seed(3000 + rep)
        model.seed(9500 + rep)
        cfg = MCTSConfig(discount=0.95, search_time_limit=0.2, c=math.sqrt(2), truncated=False, action_selection="ucb",
                         state_belief_only=False, seed=rep)
        sp = plugin_policies_for(model, level, SearchPolicyWrapper) if scen["plugins"] else None
        planner = OnDouble.initialize(model, "0", cfg, level, sp, max_sims_per_level=scen["n_sims_per_level"])
        assert planner.nesting_level == level and planner.other_agent_policies["1"].nesting_level == level - 1
        planner.reset()
        planner.update(None, obs0)
        for which in ("first", "second"):
            assert planner.get_action() in range(5)
            assert planner.step_statistics["num_sims"] == scen["n_sims_per_level"] * (level + 1)
            root = planner.traverse(planner.history)
            vis = np.array([root.children[a].visits for a in range(5)], dtype=float)
            if which == "first":          # only the top level's share of the budget simulates on the top tree
                assert vis.sum() == scen["n_sims_per_level"]
            rows[which].append(([root.children[a].value for a in range(5)], vis / vis.sum(), root.belief.size()))
            if which == "first":
                planner.update(scen["ego_action"], obs1)
                nested.append(len(planner.get_nested_history_dist({planner.history: 1.0})["1"]))
    for which in ("first", "second"):
        vals, frac, size = zip(*rows[which])
        for got, key, margin in ((vals, "value", 0.01), (frac, "visit_frac", 0.03)):
            ok, info = _close(got, scen[which][key + "_mean"], scen[which][key + "_std"], scen["reps"], margin)
            assert ok, (which, key, info)
        assert abs(np.mean(size) - scen[which]["belief_size_mean"]) < 0.25 * scen[which]["belief_size_mean"] + 2
    se = math.sqrt(scen["nested_histories_std"] ** 2 / scen["reps"] + np.std(nested) ** 2 / reps)
    assert abs(np.mean(nested) - scen["nested_histories_mean"]) < 4 * se + 0.5


def test_general_intmcp_facade(double_model):
    """step / absorbing behaviour / statistics keys of the general planner, and the routing of INTMCP.initialize."""
    from posggym_baselines_b200.planning import INTMCP, MCTSConfig, PlanningStatTracker
    from posggym_baselines_b200.planning.host_intmcp import HostINTMCP

    class OnDouble(HostINTMCP):
        def _device_initial_belief(self, init_obs, need):
            return self.model.initial_belief(self.agent_id, init_obs, need)

    model = double_model
    cfg = MCTSConfig(discount=0.95, search_time_limit=0.1, c=1.414, truncated=False, action_selection="pucb",
                     state_belief_only=False, seed=0, epsilon=0.92)
    random.seed(0)
    model.seed(0)
    planner = OnDouble.initialize(model, "0", cfg, 2, None, max_sims_per_level=12)
    state = model.sample_initial_state()
    obs = model.sample_initial_obs(state)
    planner.reset()
    rng = random.Random(0)
    for t in range(4):
        a = planner.step(obs["0"])
        assert a in range(5) and set(planner.step_statistics) == set(PlanningStatTracker.STAT_KEYS)
        assert planner.step_statistics["num_sims"] == 36 and len(planner.history) == t + 1
        ts = model.step(state, {"0": a, "1": rng.randrange(5)})
        state, obs = ts.state, ts.observations
        if ts.all_done:
            break
    assert planner.stat_tracker.get_episode()["num_sims"] == 36
    # routing: BASELINE's configuration is the device planner's, which refuses a model without the CUDA facade;
    # any other hierarchy is the general planner
    with pytest.raises(Exception):
        INTMCP.initialize(type("M", (), {"possible_agents": ("0", "1"), "action_spaces": model.action_spaces, "spec": None})(),
                          "0", cfg, 1, None)
    deep = INTMCP.initialize(model, "0", cfg, 3, None, max_sims_per_level=2)
    assert isinstance(deep, HostINTMCP) and deep.other_agent_policies["1"].other_agent_policies["0"].nesting_level == 1


@pytest.mark.parametrize("state_belief_only", [True, False])
def test_host_engine_is_the_reference_planner(double_model, state_belief_only):
    """The plugin path's engine with the Random pair (as subclasses, so that ``MCTS`` would route them to the host):
    root action values and visit shares after 192 simulations against the golden of the UNMODIFIED reference POMCP
    (tests/golden/planner_driving.json), with state-only particles and with history / policy-state particles."""
    from posggym_baselines_b200.planning import MCTSConfig, RandomOtherAgentPolicy, RandomSearchPolicy
    from posggym_baselines_b200.planning.host_search import HostSearch

    class MySearch(RandomSearchPolicy):
        pass

    class MyOther(RandomOtherAgentPolicy):
        pass

    scen = json.load(open(os.path.join(GOLDEN, "planner_driving.json")))["root_stats"][0]
    obs = _driving_obs(scen["ego_obs"])
    model, reps = double_model, 10
    vals, fracs = [], []
    for rep in range(reps):
        random.seed(50 + rep)
        model.seed(150 + rep)
        cfg = MCTSConfig(discount=0.95, search_time_limit=0.1, c=1.414, truncated=False, action_selection="ucb",
                         state_belief_only=state_belief_only, seed=rep)
        shell = _PlannerShell(model, "0", cfg, {"1": MyOther(model, "1")}, MySearch(model, "0"), scen["n_sims"])
        eng = HostSearch(shell)
        eng.update(None, obs)
        assert eng.get_action() in range(5)
        p0 = eng.root.belief.particles[0]
        assert (p0.history is None and p0.policy_state is None) if state_belief_only else (len(p0.history) == 1 and p0.policy_state == {"1": {}})
        vis = np.array([eng.root.children[a].visits for a in range(5)], dtype=float)
        vals.append([eng.root.children[a].value for a in range(5)])
        fracs.append(vis / vis.sum())
    ok, info = _close(vals, scen["value_mean"], scen["value_std"], scen["reps"], 0.015)
    assert ok, info
    assert (np.abs(np.mean(fracs, 0) - np.array(scen["visit_frac_mean"])) < 0.05).all()


def test_experiment_loop_with_belief_statistics_on_the_host_engine(double_model, tmp_path):
    """SURVEY 8(f) N3 + N4 together, on CPU: the experiment loop (``run_planning_exp``) over an agent-folding env
    wrapper, a typed planner (the host engine with ``OtherAgentMixturePolicy`` particles) and ``BeliefStatTracker``
    reading that planner's root belief -- the row has the reference's columns, the policy / action accuracies are
    real probabilities and the belief concentrates on the other agent's true type over an episode."""
    import csv

    from posggym_baselines_b200.envs.agent_env import AgentEnvWrapper, UniformOtherAgentFn
    from posggym_baselines_b200.envs.env import Env
    from posggym_baselines_b200.exp import episode_results_heads, run_planning_exp
    from posggym_baselines_b200.planning import (MCTSConfig, OtherAgentMixturePolicy, PlanningStatTracker,
                                                 SearchPolicyWrapper)
    from posggym_baselines_b200.planning.host_search import HostSearch
    from tests.potmmcp_policies import BiasedPolicy, other_agent_types

    model = double_model
    cfg = MCTSConfig(discount=0.95, search_time_limit=0.05, c=1.25, truncated=True, action_selection="pucb",
                     state_belief_only=False, seed=0)

    class HostPlanner(_PlannerShell):
        """``MCTS`` in host mode (planning/mcts.py: reset / step / root) without the device search around it."""

        def __init__(self):
            super().__init__(model, "0", cfg, {"1": OtherAgentMixturePolicy(model, "1", other_agent_types(model))},
                             SearchPolicyWrapper(BiasedPolicy(model, "0", "fast", 1, 4.0, value=0.1)), 48)
            self.stat_tracker = PlanningStatTracker(self)
            self.engine = HostSearch(self)
            self._last = None

        root = property(lambda self: self.engine.root)

        def reset(self):
            self.stat_tracker.reset_episode()
            self.engine.reset()
            self._last = None

        def step(self, obs):
            self.engine.update(self._last, obs)
            self._last = self.engine.get_action()
            self.stat_tracker.step()
            return self._last

    random.seed(3)
    model.seed(3)
    env = AgentEnvWrapper(Env(model), UniformOtherAgentFn({"1": ["pusher"]}, lambda pid, m, i: other_agent_types(m, i)[pid]))
    stats = ["policy", "action", "state", "history"]
    out = tmp_path / "episode_results.csv"
    rows = run_planning_exp(env, HostPlanner(), "0", None, 2, str(out), belief_stats_to_track=stats, seed=5)
    with open(out) as f:
        got = list(csv.DictReader(f))
    assert list(got[0].keys()) == episode_results_heads([f"belief_{k}_acc" for k in stats]) and len(got) == 2
    for r in rows:
        assert 0.3 < r["belief_policy_acc"] <= 1.0          # two types, the true one is 'pusher': above chance
        assert 0.0 < r["belief_action_acc"] <= 1.0 and 0.0 <= r["belief_state_acc"] <= 1.0
        assert r["belief_history_acc"] * r["len"] <= 1.0 + 1e-9 and r["num_sims"] == 48
