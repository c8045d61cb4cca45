"""The drop-in boundary from both sides:

* the UNMODIFIED reference planner (imported through oracle.shims from /root/reference when that tree is
  present; on the GPU box it is not, and those tests skip) driving the GPU model facade produces, call for
  call, what it produces on the CPU oracle model under the same RNG seeds;
* arbitrary ``SearchPolicy`` / ``OtherAgentPolicy`` plugins run through the B200 planner's host path
  (planning/host_search.py) -- the reference's own tests/planning/test_mcts.py scenario included.
"""
import math
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HAVE_REF = os.path.isdir("/root/reference/posggym_baselines")


class FakeClock:
    """``time`` stand-in for the reference planner: the search budget becomes a fixed number of clock reads."""

    def __init__(self, tick):
        self.now, self.tick = 0.0, tick

    def time(self):
        self.now += self.tick
        return self.now


def _run_reference_planner(model, env_cls, seed, clock_tick, max_steps=12):
    import posggym_baselines.planning.mcts as ref_mcts
    from posggym_baselines.planning import POMCP, MCTSConfig, RandomSearchPolicy

    ref_mcts.time = FakeClock(clock_tick)
    random.seed(seed)
    np.random.seed(seed)
    cfg = MCTSConfig(discount=0.95, search_time_limit=0.1, c=1.414, truncated=False, action_selection="ucb",
                     state_belief_only=True, seed=seed)
    env = env_cls(model)
    obs, _ = env.reset(seed=seed)
    planner = POMCP(model, "0", cfg, search_policy=RandomSearchPolicy(model, "0"))
    planner.reset()
    rng = random.Random(seed + 1)
    trace = []
    for _ in range(max_steps):
        a = planner.step(obs["0"])
        b = rng.randrange(5)
        obs, rew, term, trunc, done, _ = env.step({"0": a, "1": b})
        trace.append((a, b, planner.step_statistics["num_sims"], obs["0"], obs["1"], rew["0"], rew["1"], done))
        if done:
            break
    planner.close()
    return trace


@pytest.mark.skipif(not HAVE_REF, reason="needs the reference tree (build container)")
@pytest.mark.parametrize("env_id", ["Driving-v1", "PredatorPrey-v0", "LevelBasedForaging-v3"])
def test_unmodified_reference_planner_on_the_gpu_facade(oracle_env, env_id):
    """Same seeds for ``random``, ``model.rng``, the action spaces and the planner; a fake clock fixes the
    number of simulations.  Every chosen action, simulation count, observation and reward of the episode
    prefix is identical -- i.e. the facade is the oracle model as far as the reference planner can tell
    (states / observations hashable and comparable, rewards bit-equal, RNG draws in the same order)."""
    import time as real_time

    import posggym_baselines.planning.mcts as ref_mcts
    from oracle.driving import DrivingModel as OracleDriving
    from oracle.lbf import LevelBasedForagingModel as OracleLBF
    from oracle.predator_prey import PredatorPreyModel as OraclePP
    from posggym_baselines_b200.envs.driving import DrivingModel
    from posggym_baselines_b200.envs.env import Env
    from posggym_baselines_b200.envs.lbf import LBFModel
    from posggym_baselines_b200.envs.predator_prey import PredatorPreyModel

    gpu_cls, cpu_cls = {"Driving-v1": (DrivingModel, OracleDriving), "PredatorPrey-v0": (PredatorPreyModel, OraclePP),
                        "LevelBasedForaging-v3": (LBFModel, OracleLBF)}[env_id]
    try:
        results = []
        for cls in (gpu_cls, cpu_cls):
            m = cls()
            if m.spec is None:
                m.spec = type("S", (), {"max_episode_steps": 50, "id": env_id})()
            results.append(_run_reference_planner(m, Env, seed=7, clock_tick=0.0004, max_steps=6))
    finally:
        ref_mcts.time = real_time
    gpu, cpu = results
    assert len(gpu) == len(cpu) and len(gpu) >= 3
    for g, c in zip(gpu, cpu):
        assert g[:3] == c[:3], (g[:3], c[:3])                  # ego action, other action, simulations
        assert g[5:] == c[5:]                                    # rewards (bit-equal floats), done
        assert _plain(g[3]) == _plain(c[3]) and _plain(g[4]) == _plain(c[4])
    assert all(s[2] > 10 for s in gpu)


def _plain(x):
    return tuple(_plain(v) for v in x) if isinstance(x, (tuple, list)) else int(x)


@pytest.mark.parametrize("env_id", ["Driving-v1", "PredatorPrey-v0", "LevelBasedForaging-v3"])
def test_restated_reference_planner_on_the_gpu_facade(oracle_env, env_id):
    """The same check where the reference tree is absent (the GPU box): oracle/pomcp.py -- the scalar
    restatement of the reference planner, pinned to the unmodified planner by tests/test_oracle_planner.py --
    runs an episode prefix on the GPU facade and on the CPU oracle model with the same seeds and a fixed
    simulation count; every action, observation and reward is identical, so the facade consumes ``model.rng``
    and the action spaces exactly like the oracle model and returns equal, hashable values."""
    from oracle.driving import DrivingModel as OracleDriving
    from oracle.lbf import LevelBasedForagingModel as OracleLBF
    from oracle.pomcp import OraclePOMCP
    from oracle.predator_prey import PredatorPreyModel as OraclePP
    from posggym_baselines_b200.envs.driving import DrivingModel
    from posggym_baselines_b200.envs.env import Env
    from posggym_baselines_b200.envs.lbf import LBFModel
    from posggym_baselines_b200.envs.predator_prey import PredatorPreyModel

    gpu_cls, cpu_cls = {"Driving-v1": (DrivingModel, OracleDriving), "PredatorPrey-v0": (PredatorPreyModel, OraclePP),
                        "LevelBasedForaging-v3": (LBFModel, OracleLBF)}[env_id]
    traces = []
    for cls in (gpu_cls, cpu_cls):
        m = cls()
        if m.spec is None:
            m.spec = type("S", (), {"max_episode_steps": 50, "id": env_id})()
        random.seed(11)
        env = Env(m)
        obs, _ = env.reset(seed=11)
        n_act = m.action_spaces["0"].n
        planner = OraclePOMCP(m, "0", 0.95, 0.1, 1.414, seed=11)
        planner.reset()
        rng, trace, last = random.Random(12), [], None
        for _ in range(6):
            planner.update(last, obs["0"])
            last = planner.get_action(max_sims=40)
            b = rng.randrange(n_act)
            obs, rew, term, trunc, done, _ = env.step({"0": last, "1": b})
            trace.append((last, b, _plain(obs["0"]), _plain(obs["1"]), rew["0"], rew["1"], done))
            if done or planner.root.is_absorbing:
                break
        traces.append(trace)
    assert traces[0] == traces[1] and len(traces[0]) >= 3


# ---------------------------------------------------------------- plugins through the B200 planner
def _plugin_classes():
    from posggym_baselines_b200.planning import OtherAgentPolicy, SearchPolicy

    class CountingSearchPolicy(SearchPolicy):
        """A stateful plugin: remembers how many steps it has seen, prefers ACCELERATE early on, and
        offers a value estimate."""

        def __init__(self, model, agent_id):
            super().__init__(model, agent_id, "counting")
            self.n = model.action_spaces[agent_id].n
            self.calls = 0

        def get_initial_state(self):
            return {"k": 0, "last_obs": None}

        def get_next_state(self, action, obs, state):
            self.calls += 1
            return {"k": state["k"] + 1, "last_obs": obs}

        def sample_action(self, state):
            return random.choices(range(self.n), weights=list(self.get_pi(state).values()))[0]

        def get_pi(self, state):
            w = [1.0] * self.n
            w[1] += 3.0 / (1 + state["k"])
            z = sum(w)
            return {a: w[a] / z for a in range(self.n)}

        def get_value(self, state):
            return 0.01 * state["k"]

    class StickyOther(OtherAgentPolicy):
        """Repeats its previous action with probability 0.8 (needs its own state per particle)."""

        def __init__(self, model, agent_id):
            super().__init__(model, agent_id)
            self.n = model.action_spaces[agent_id].n

        def sample_initial_state(self):
            return {"last": None}

        def get_next_state(self, action, obs, state):
            return {"last": action}

        def sample_action(self, state):
            if state["last"] is not None and random.random() < 0.8:
                return state["last"]
            return random.randrange(self.n)

        def get_pi(self, state):
            return {a: 1.0 / self.n for a in range(self.n)}

    return CountingSearchPolicy, StickyOther


@pytest.mark.parametrize("truncated", [True, False])
def test_stateful_plugins_run_through_the_host_path(truncated):
    from posggym_baselines_b200.envs.driving import DrivingModel
    from posggym_baselines_b200.envs.env import Env
    from posggym_baselines_b200.planning import MCTS, MCTSConfig, PlanningStatTracker

    Search, Other = _plugin_classes()
    model = DrivingModel()
    env = Env(model)
    cfg = MCTSConfig(discount=0.95, search_time_limit=0.05, c=1.414, truncated=truncated, action_selection="pucb",
                     state_belief_only=False, seed=3)
    sp = Search(model, "0")
    planner = MCTS(model, "0", cfg, {"1": Other(model, "1")}, sp, max_sims=24 if truncated else 6)
    assert planner._host is not None
    obs, _ = env.reset(seed=3)
    planner.reset()
    random.seed(0)
    for step in range(4):
        a = planner.step(obs["0"])
        assert a in range(5)
        st = planner.step_statistics
        assert set(st) == set(PlanningStatTracker.STAT_KEYS)
        assert st["num_sims"] == (24 if truncated else 6) and st["policy_calls"] > 0
        root = planner.root
        assert root.t == step + 1 and root.belief.size() >= cfg.num_particles + cfg.extra_particles
        p0 = root.belief.particles[0]
        assert len(p0.history) == step + 1 and set(p0.policy_state) == {"1"} and p0.t == step + 1
        assert set(p0.policy_state["1"]) == {"last"}
        obs, _, _, _, done, _ = env.step({"0": a, "1": random.randrange(5)})
        if done:
            break
    assert sp.calls > 0
    planner.close()


def test_reference_mcts_test_scenario():
    """tests/planning/test_mcts.py:13-27 of the reference: MCTS with the Random pair, epsilon 0.92
    (depth_limit 2), state_belief_only=False -- runs on the device search."""
    from posggym_baselines_b200.envs.driving import DrivingModel
    from posggym_baselines_b200.envs.env import Env
    from posggym_baselines_b200.planning import MCTS, MCTSConfig, RandomOtherAgentPolicy, RandomSearchPolicy

    model = DrivingModel()
    env = Env(model)
    cfg = MCTSConfig(discount=0.95, search_time_limit=0.1, c=math.sqrt(2), truncated=False, action_selection="ucb",
                     pucb_exploration_fraction=0.25, known_bounds=None, step_limit=None, epsilon=0.92, seed=0,
                     state_belief_only=False)
    assert cfg.depth_limit == 2
    planner = MCTS(model, "0", cfg, other_agent_policies={"1": RandomOtherAgentPolicy(model, "1")},
                   search_policy=RandomSearchPolicy(model, "0"), node_capacity=1 << 22, log_capacity=1 << 24)
    assert planner._host is None
    obs, _ = env.reset(seed=0)
    planner.reset()
    rng, done, steps = random.Random(0), False, 0
    while not done:
        a = planner.step(obs["0"])
        obs, _, _, _, done, _ = env.step({"0": a, "1": rng.randrange(5)})
        steps += 1
    assert steps <= 50
    planner.close()


def test_random_pair_subclasses_match_the_device_search_statistically():
    """The host path is the same algorithm as the device path: a SUBCLASS of the Random pair (not recognised by
    type, so it goes through host_search.py) gives root action values within 4 standard errors + 0.03 of the
    device search after 192 simulations from the same first observation."""
    import json

    from posggym_baselines_b200.envs.driving import DrivingModel
    from posggym_baselines_b200.planning import MCTS, MCTSConfig, RandomOtherAgentPolicy, RandomSearchPolicy
    from tests.conftest import GOLDEN

    class MySearch(RandomSearchPolicy):
        pass

    class MyOther(RandomOtherAgentPolicy):
        pass

    scen = json.load(open(os.path.join(GOLDEN, "planner_driving.json")))["root_stats"][0]
    f = scen["ego_obs"]
    obs = (tuple(f[:15]), f[15], (f[16], f[17]), (f[18], f[19]), f[20], f[21])
    model = DrivingModel()
    vals = []
    for rep in range(6):
        cfg = MCTSConfig(discount=0.95, search_time_limit=0.1, c=1.414, truncated=False, action_selection="ucb",
                         state_belief_only=True, seed=rep)
        random.seed(rep)
        model.seed(100 + rep)
        p = MCTS(model, "0", cfg, {"1": MyOther(model, "1")}, MySearch(model, "0"), max_sims=192)
        assert p._host is not None
        p.reset()
        p.update(None, obs)
        p.get_action()
        vals.append([p.root.children[a].value for a in range(5)])
        p.close()
    vals = np.array(vals)
    ref_mean, ref_std = np.array(scen["value_mean"]), np.array(scen["value_std"])
    se = np.sqrt(ref_std**2 / scen["reps"] + vals.std(0) ** 2 / len(vals))
    assert (np.abs(vals.mean(0) - ref_mean) < 4 * se + 0.03).all(), (vals.mean(0), ref_mean, se)


def test_ppo_lstm_search_policy_combined_planner():
    """SURVEY 8(f) N3, first slice -- the COMBINED planner of baseline_exps/run_combined_exps.py:48-91: MCTS with
    a trained PPOLSTMModel as search policy (PUCB prior, value-function leaf evaluation, truncated=True).  The
    network is evaluated on the GPU (tensor-core policy kernel); its prior / value / LSTM state agree with the
    fp32 PyTorch module within 5e-3, and the planner runs an episode prefix through the host path."""
    import torch

    from posggym_baselines_b200.envs.driving import DrivingModel
    from posggym_baselines_b200.envs.env import Env
    from posggym_baselines_b200.planning import MCTS, MCTSConfig, PPOLSTMSearchPolicy, RandomOtherAgentPolicy
    from tests.test_policy_gpu import RefNet

    torch.manual_seed(0)
    model = DrivingModel()
    dev = model.batched.device
    net = RefNet(124, 5).to(dev)
    sp = PPOLSTMSearchPolicy(model, "0", "BR_test_seed0", net.state_dict())
    env = Env(model)
    obs, _ = env.reset(seed=2)
    # one-hot flattening of the observation tuple = the layout the reference's FlattenProcessor produces
    leaves = [(0, 4)] * 15 + [(0, 4), (0, 14), (0, 14), (0, 14), (0, 14), (0, 2), (0, 2)]
    flat, off = torch.zeros(124, device=dev), 0
    vals = list(obs["0"][0]) + [obs["0"][1], *obs["0"][2], *obs["0"][3], obs["0"][4], obs["0"][5]]
    for (_, n), v in zip(leaves, vals):
        flat[off + v] = 1.0
        off += n
    assert off == 124
    st = sp.get_next_state(None, obs["0"], sp.get_initial_state())
    with torch.no_grad():
        logits, _, _, val, (h1, c1) = net.get_action_and_value(flat.unsqueeze(0), (torch.zeros(1, 1, 64, device=dev),) * 2,
                                                               torch.zeros(1, device=dev), torch.zeros(1, dtype=torch.int64, device=dev))
    ref_pi = torch.softmax(logits[0], -1).cpu().numpy()
    assert np.abs(np.array([st["pi"][a] for a in range(5)]) - ref_pi).max() < 5e-3
    assert abs(st["value"] - val.item()) < 5e-3 and (st["lstm_state"][0] - h1[0].cpu()).abs().max().item() < 5e-3
    cfg = MCTSConfig(discount=0.99, search_time_limit=0.1, c=1.25, truncated=True, action_selection="pucb",
                     state_belief_only=False, seed=1)
    planner = MCTS(model, "0", cfg, {"1": RandomOtherAgentPolicy(model, "1")}, sp)
    assert planner._host is not None
    planner.reset()
    rng, sims = random.Random(0), []
    for _ in range(3):
        a = planner.step(obs["0"])
        assert a in range(5)
        sims.append(planner.step_statistics["num_sims"])
        obs, _, _, _, done, _ = env.step({"0": a, "1": rng.randrange(5)})
        if done:
            break
    assert min(sims) >= 5, sims        # the reference's COMBINED planner does ~37 simulations per 0.1 s on Driving
    print("COMBINED planner simulations per 0.1 s decision:", sims)
    planner.close()


def _driving_obs(f):
    return (tuple(f[:15]), f[15], (f[16], f[17]), (f[18], f[19]), f[20], f[21])


@pytest.mark.parametrize("truncated", [False, True])
def test_potmmcp_matches_the_unmodified_reference_planner(truncated):
    """POTMMCP (typed other agent = ``OtherAgentMixturePolicy``, meta-policy over the planning
    agent's policies) through the host engine with the model on the device.  tests/golden/potmmcp_driving.json
    holds root statistics of the UNMODIFIED reference POTMMCP with the same policy objects
    (tests/potmmcp_policies.py) over 64 repetitions, for the first decision and for the one after
    ``update(action, obs)``; action values, visit shares, the root prior after the search's moving-average updates
    and the share of 'pusher' particles in the updated belief agree within 4 standard errors plus a small margin."""
    import json

    from posggym_baselines_b200.envs.driving import DrivingModel
    from posggym_baselines_b200.planning import POTMMCP, MCTSConfig, OtherAgentMixturePolicy, POTMMCPMetaPolicy
    from tests.conftest import GOLDEN
    from tests.potmmcp_policies import META_POLICY, ego_policies, other_agent_types

    gold = json.load(open(os.path.join(GOLDEN, "potmmcp_driving.json")))
    scen = next(s for s in gold["scenarios"] if s["truncated"] == truncated)
    obs0, obs1 = _driving_obs(scen["obs0"]), _driving_obs(scen["obs1"])
    model = DrivingModel()
    reps = 8
    rows = {"first": [], "second": []}
    shares = []
    for rep in range(reps):
        random.seed(1000 + rep)
        model.seed(7000 + rep)
        cfg = MCTSConfig(discount=0.95, search_time_limit=0.1, c=1.25, truncated=truncated, action_selection="pucb",
                         pucb_exploration_fraction=0.5, state_belief_only=False, seed=rep)
        ego = ego_policies(model)
        for pi in ego.values():
            pi.value = 0.05 if truncated else None
        planner = POTMMCP(model, "0", cfg, {"1": OtherAgentMixturePolicy(model, "1", other_agent_types(model))},
                          POTMMCPMetaPolicy(model, "0", ego, META_POLICY), max_sims=scen["n_sims"])
        planner.reset()
        planner.update(None, obs0)
        for which, nxt in (("first", obs1), ("second", None)):
            assert planner.get_action() in range(5)
            root = planner.root
            assert planner.step_statistics["num_sims"] == scen["n_sims"] and root.visits >= scen["n_sims"]
            vis = np.array([root.children[a].visits for a in range(5)], dtype=float)
            rows[which].append(([root.children[a].value for a in range(5)], vis / vis.sum(),
                                [root.action_probs[a] for a in range(5)]))
            if nxt is not None:
                planner.update(scen["ego_action"], nxt)
                parts = planner.root.belief.particles
                assert len(parts) >= cfg.num_particles + cfg.extra_particles
                assert all(len(p.history) == 2 and p.t == 2 for p in parts)
                shares.append(np.mean([p.policy_state["1"]["policy_id"] == "pusher" for p in parts]))
        planner.close()
    for which in ("first", "second"):
        g = scen[which]
        vals, frac, prior = (np.array(x) for x in zip(*rows[which]))
        for got, key, margin in ((vals, "value", 0.03), (frac, "visit_frac", 0.05), (prior, "prior", 0.02)):
            mean, std = np.array(g[key + "_mean"]), np.array(g[key + "_std"])
            se = np.sqrt(std**2 / scen["reps"] + got.std(0) ** 2 / reps)
            assert (np.abs(got.mean(0) - mean) < 4 * se + margin).all(), (which, key, got.mean(0), mean, se)
        assert prior[:, 1].std() > 0.4 * g["prior_std"][1]      # the root prior moves with the search, as in the reference
    se = math.sqrt(scen["pusher_share_std"] ** 2 / scen["reps"] + np.std(shares) ** 2 / reps)
    assert abs(np.mean(shares) - scen["pusher_share_mean"]) < 4 * se + 0.05, (np.mean(shares), scen["pusher_share_mean"])


def test_ipomcp_with_typed_other_agent_plays_an_episode():
    """IPOMCP (planning/ipomcp.py) = MCTS over history / policy-state particles: with a mixture of other-agent types
    and a wrapped ``posggym.agents``-style search policy it runs an episode through the host engine."""
    from posggym_baselines_b200.envs.driving import DrivingModel
    from posggym_baselines_b200.envs.env import Env
    from posggym_baselines_b200.planning import IPOMCP, MCTSConfig, OtherAgentMixturePolicy, SearchPolicyWrapper
    from tests.potmmcp_policies import BiasedPolicy, other_agent_types

    model = DrivingModel()
    env = Env(model)
    cfg = MCTSConfig(discount=0.95, search_time_limit=0.05, c=1.25, truncated=True, action_selection="pucb",
                     state_belief_only=False, seed=4)
    planner = IPOMCP(model, "0", cfg, {"1": OtherAgentMixturePolicy(model, "1", other_agent_types(model))},
                     SearchPolicyWrapper(BiasedPolicy(model, "0", "fast", 1, 4.0, value=0.1)), max_sims=64)
    assert planner._host is not None
    obs, _ = env.reset(seed=4)
    planner.reset()
    rng = random.Random(4)
    for step in range(5):
        a = planner.step(obs["0"])
        assert a in range(5) and planner.step_statistics["num_sims"] == 64
        ids = {p.policy_state["1"]["policy_id"] for p in planner.root.belief.particles}
        assert ids <= {"pusher", "idler"} and ids
        obs, _, _, _, done, _ = env.step({"0": a, "1": rng.randrange(5)})
        if done:
            break
    planner.close()
