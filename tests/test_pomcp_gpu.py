"""The CUDA POMCP planner: API smoke, sequential-vs-wave consistency and STATISTICAL
parity with the unmodified reference planner (golden statistics generated on the CPU
oracle by tests/golden/make_planner_golden.py).  Tolerances are stated per test."""
import json
import math
import os
import random

import numpy as np
import pytest
import torch

from tests.conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(GOLDEN, "planner_driving.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def model():
    from posggym_baselines_b200.envs.driving import DrivingModel

    return DrivingModel()


def make_planner(model, seed=0, **kw):
    from posggym_baselines_b200.planning import POMCP, MCTSConfig, RandomSearchPolicy

    opts = {k: kw.pop(k) for k in list(kw) if k in ("max_sims", "wave_size", "max_wave", "node_capacity", "log_capacity",
                                                    "trees_per_gpu", "_forest_index")}
    if opts.get("max_sims") is not None:   # fixed small budgets: a small arena is plenty
        opts.setdefault("node_capacity", 1 << 18)
        opts.setdefault("log_capacity", 1 << 21)
    cfg = dict(discount=0.95, search_time_limit=0.1, c=1.414, truncated=False, action_selection="ucb", seed=seed)
    cfg.update(kw)
    return POMCP(model, "0", MCTSConfig(**cfg), RandomSearchPolicy(model, "0"), **opts)


def fields_to_obs(f):
    return (tuple(f[:15]), f[15], (f[16], f[17]), (f[18], f[19]), f[20], f[21])


def test_readme_loop(model):
    """README.md:52-88 -- reset/step loop, statistics keys, wall-clock budget."""
    from posggym_baselines_b200.envs.env import Env
    from posggym_baselines_b200.planning import PlanningStatTracker

    env = Env(model)
    planner = make_planner(model, seed=1)   # README config: state_belief_only defaults False -> POMCP copies it
    assert planner.config.state_belief_only and planner.config.num_particles == 10 and planner.config.depth_limit == 90
    obs, _ = env.reset(seed=1)
    planner.reset()
    rng = random.Random(1)
    done, steps = False, 0
    while not done:
        a = planner.step(obs["0"])
        assert a in range(5)
        st = planner.step_statistics
        assert set(st) == set(PlanningStatTracker.STAT_KEYS)
        # once the ego vehicle is done the root is absorbing: no search (mcts.py:270-272)
        if not math.isnan(st["num_sims"]) and not planner.root.is_absorbing:
            assert st["num_sims"] >= 1000 and 0.1 <= st["search_time"] < 0.5
        obs, rew, term, trunc, done, _ = env.step({"0": a, "1": rng.randrange(5)})
        steps += 1
    assert steps <= 50
    ep = planner.stat_tracker.get_episode()
    assert "num_sims" in ep
    planner.close()


def run_root(model, scen, n_sims, reps, **opts):
    vals, vis = [], []
    for rep in range(reps):
        p = make_planner(model, seed=rep, max_sims=n_sims, **opts)
        p.reset()
        p.update(None, fields_to_obs(scen["ego_obs"]))
        p.get_action()
        st = p._last_root_stats
        vals.append(np.where(st["done"] > 0, st["total"] / np.maximum(st["done"], 1), 0.0))
        vis.append(st["done"] / st["done"].sum())
    return np.array(vals), np.array(vis)


@pytest.mark.parametrize("wave", [1, 32])
def test_root_statistics_match_reference(model, gold, wave):
    """Root action values and visit shares after N simulations from the same first
    observation: mean over independent planners within 4 standard errors (+0.02 abs) of the
    reference planner's means; the best action agrees."""
    for scen in gold["root_stats"]:
        reps = 60
        vals, vis = run_root(model, scen, scen["n_sims"], reps, wave_size=wave)
        ref_mean, ref_std = np.array(scen["value_mean"]), np.array(scen["value_std"])
        se = np.sqrt(ref_std**2 / scen["reps"] + vals.std(0) ** 2 / reps)
        assert (np.abs(vals.mean(0) - ref_mean) < 4 * se + 0.02).all(), (vals.mean(0), ref_mean, se)
        assert (np.abs(vis.mean(0) - np.array(scen["visit_frac_mean"])) < 0.05).all(), (vis.mean(0), scen["visit_frac_mean"])
        order = np.argsort(ref_mean)
        if ref_mean[order[-1]] - ref_mean[order[-2]] > 4 * se.max() + 0.02:   # only when the reference's best is clear
            assert int(np.argmax(vals.mean(0))) == int(order[-1])


def test_forest_tree_equals_solo_tree(model, gold):
    """trees_per_gpu = K (root parallelisation inside the GPU, posg_search_forest): tree k of the forest IS the
    single-tree search with tree k's Philox sub-sequences -- with wave size 1 (deterministic) its root statistics
    are exactly those of a solo planner given the same counters; the forest's decision statistics are their sum."""
    scen = gold["root_stats"][0]
    obs = fields_to_obs(scen["ego_obs"])
    K, n = 3, 512
    f = make_planner(model, seed=7, max_sims=n, wave_size=1, trees_per_gpu=K)
    f.reset(); f.update(None, obs); f.get_action()
    assert f.step_statistics["num_sims"] == K * n and f.search_info["trees"] == K
    trees = [f] + f._siblings
    total_done = np.zeros(5)
    for k in range(K):
        solo = make_planner(model, seed=7, max_sims=n, wave_size=1, _forest_index=k)
        solo._sim_base = k << 32
        solo.reset(); solo.update(None, obs); solo.get_action()
        a, b_ = trees[k]._last_root_stats, solo._last_root_stats
        assert (a["done"] == b_["done"]).all() and (a["total"] == b_["total"]).all(), (k, a["done"], b_["done"])
        assert a["node_count"] == b_["node_count"]
        total_done += a["done"]
        solo.close()
    assert (f._merged_root[0] == total_done).all()
    assert not (trees[0]._last_root_stats["total"] == trees[1]._last_root_stats["total"]).all()   # different streams
    f.close()


def test_forest_root_statistics_match_reference(model, gold):
    """Merged root statistics of a 4-tree forest, every tree at the reference's simulation count: the action
    values are within 4 standard errors (+0.02) of the reference planner's means (each tree is an independent
    sample of the reference's search, so the visit-weighted mean has the same expectation)."""
    for scen in gold["root_stats"]:
        reps = 20
        vals = []
        for rep in range(reps):
            p = make_planner(model, seed=100 + rep, max_sims=scen["n_sims"], wave_size=1, trees_per_gpu=4)
            p.reset(); p.update(None, fields_to_obs(scen["ego_obs"])); p.get_action()
            vals.append(p.root_action_values())
            p.close()
        vals = np.array(vals)
        ref_mean, ref_std = np.array(scen["value_mean"]), np.array(scen["value_std"])
        se = np.sqrt(ref_std**2 / scen["reps"] + vals.std(0) ** 2 / reps)
        assert (np.abs(vals.mean(0) - ref_mean) < 4 * se + 0.02).all(), (vals.mean(0), ref_mean, se)


def test_forest_episode_and_throughput(model, gold):
    """An episode with 16 trees on one GPU under the time budget: the re-root / belief update runs on every tree,
    nothing overflows, and the decision sees (about) 16 x the single tree's simulations."""
    import random as _random
    from posggym_baselines_b200.envs.env import Env

    planner = make_planner(model, seed=9, trees_per_gpu=16)
    env = Env(model)
    rng = _random.Random(9)
    obs, _ = env.reset(seed=9)
    planner.reset()
    sims, per_tree = [], []
    for d in range(6):
        a = planner.step(obs["0"])
        assert a in range(5)
        if planner.root.is_absorbing:
            break
        info = planner.search_info
        assert info["overflow_share"] == 0.0 and info["trees"] == 16
        if d > 0:
            sims.append(info["num_sims"]); per_tree.append(info["sims_per_tree"])
        for t in planner._siblings:
            assert t._root_t == planner._root_t and t.root.belief.size() >= planner.config.num_particles
        obs, _, _, _, done, _ = env.step({"0": a, "1": rng.randrange(5)})
        if done:
            break
    assert len(sims) >= 3
    assert min(per_tree) >= 300 * gold["episodes"]["mean_sims_per_decision"], per_tree
    assert sum(sims) / len(sims) >= 8000 * gold["episodes"]["mean_sims_per_decision"], sims
    planner.close()


def test_forest_episodes_not_worse_than_one_tree(model):
    """Episode outcomes with 8 trees per GPU against one tree at the same time budget (0.02 s per decision, the
    planner's own wave schedule in both): merging independent searches at the root must not play worse -- success
    rate not lower by more than 3 binomial standard errors + 0.05, returns not told apart by a KS test at p > 0.001."""
    from posggym_baselines_b200.envs.env import Env

    env = Env(model)

    def play(n, **opts):
        wins, rets = 0, []
        for ep in range(n):
            planner = make_planner(model, seed=ep, search_time_limit=0.02, node_capacity=1 << 24, log_capacity=1 << 26, **opts)
            rng = random.Random(30_000 + ep)
            obs, _ = env.reset(seed=2000 + ep)
            planner.reset()
            done, R = False, 0.0
            while not done:
                a = planner.step(obs["0"])
                obs, rew, term, trunc, done, _ = env.step({"0": a, "1": rng.randrange(5)})
                R += rew["0"]
            wins += env.state[0].dest_reached
            rets.append(R)
            planner.close()
        return wins / n, np.array(rets)

    n = 24
    p_one, r_one = play(n)
    p_forest, r_forest = play(n, trees_per_gpu=8)
    se = math.sqrt(p_one * (1 - p_one) / n + p_forest * (1 - p_forest) / n)
    assert p_forest > p_one - 3 * se - 0.05, (p_forest, p_one)
    d, pv = ks_2samp(r_forest, r_one)
    assert pv > 0.001, (d, pv, r_forest.mean(), r_one.mean())


def ks_2samp(a, b):
    a, b = np.sort(a), np.sort(b)
    allv = np.concatenate([a, b])
    cdf_a = np.searchsorted(a, allv, side="right") / len(a)
    cdf_b = np.searchsorted(b, allv, side="right") / len(b)
    d = np.abs(cdf_a - cdf_b).max()
    en = math.sqrt(len(a) * len(b) / (len(a) + len(b)))
    lam = (en + 0.12 + 0.11 / en) * d
    p = 2 * sum((-1) ** (k - 1) * math.exp(-2 * k * k * lam * lam) for k in range(1, 101))
    return d, max(0.0, min(1.0, p))


def test_rollout_return_distribution_ks(model, gold):
    """Kolmogorov-Smirnov two-sample test of the fused rollout kernel's returns against
    MCTS._rollout of the reference (its own RNG streams): not rejected at p > 0.001."""
    from posggym_baselines_b200.envs.driving import VehicleState

    b = model.batched
    for scen in gold["rollouts"]:
        st = tuple(VehicleState((f[0], f[1]), f[2], f[3], (f[4], f[5]), f[6], f[7], f[8], f[9]) for f in scen["state"])
        packed = torch.from_numpy(np.repeat(b.pack_state([st]), 50_000, axis=0)).to(b.device)
        max_steps = min(scen["depth_limit"] + 1, scen["step_limit"] - 1 + 1)
        ret, _, _ = b.rollout(packed, 0, max_steps, scen["discount"], seed=123)
        d, p = ks_2samp(ret.cpu().numpy(), np.array(scen["returns_sorted"]))
        assert p > 0.001, (d, p)


def test_episode_returns_match_reference(model, gold):
    """POMCP vs a uniform-random other agent at (about) the reference's simulation count
    per decision: success rate within 3 binomial standard errors (+0.05)."""
    from posggym_baselines_b200.envs.env import Env

    ref = gold["episodes"]
    n_sims = int(round(ref["mean_sims_per_decision"]))
    env = Env(model)
    wins, rets, n = 0, [], 60
    for ep in range(n):
        planner = make_planner(model, seed=ep, max_sims=n_sims, wave_size=1)
        rng = random.Random(10_000 + ep)
        obs, _ = env.reset(seed=ep)
        planner.reset()
        done, R = False, 0.0
        while not done:
            a = planner.step(obs["0"])
            obs, rew, term, trunc, done, _ = env.step({"0": a, "1": rng.randrange(5)})
            R += rew["0"]
        wins += env.state[0].dest_reached
        rets.append(R)
        planner.close()
    p_ref = ref["outcomes"].count("WIN") / ref["n_episodes"]
    p = wins / n
    se = math.sqrt(p_ref * (1 - p_ref) / ref["n_episodes"] + p * (1 - p) / n)
    assert abs(p - p_ref) < 3 * se + 0.05, (p, p_ref)
    assert abs(np.mean(rets) - np.mean(ref["returns"])) < 3 * math.sqrt(
        np.var(rets) / n + np.var(ref["returns"]) / ref["n_episodes"]) + 0.05


def test_throughput_mode_beats_reference_by_1000x(model, gold):
    """>= 1000x the reference's simulations per 0.1 s decision (north_star), measured like bench.py: the mean
    over the decisions of an episode after the first (which pays for the allocations).  The first decisions run
    at ~7e4 simulations (one- and two-warp waves on a young tree); once the kept subtree carries visits the
    schedule allows larger waves."""
    import random as _random

    planner = make_planner(model, seed=5)
    from posggym_baselines_b200.envs.env import Env

    env = Env(model)
    rng = _random.Random(5)
    obs, _ = env.reset(seed=5)
    planner.reset()
    sims = []
    for d in range(9):
        a = planner.step(obs["0"])
        if d > 0:
            sims.append(planner.step_statistics["num_sims"])
        assert planner.search_info["overflow_share"] == 0.0
        obs, _, _, _, done, _ = env.step({"0": a, "1": rng.randrange(5)})
        if done:
            break
    assert len(sims) >= 4
    assert min(sims) >= 500 * gold["episodes"]["mean_sims_per_decision"], sims
    assert sum(sims) / len(sims) >= 1000 * gold["episodes"]["mean_sims_per_decision"], sims


def test_reroot_keeps_the_subtree(model):
    """MCTS._update (mcts.py:229-263) re-roots at the (action, observation) child and keeps its
    subtree: after posg_tree_reroot the new root (node 0) has the child's statistics, the
    particles logged at it are the new root belief, and deeper nodes are still reachable."""
    import numpy as np

    from posggym_baselines_b200 import _lib
    from posggym_baselines_b200.envs.env import Env

    env = Env(model)
    p = make_planner(model, seed=3, max_sims=6000, wave_size=32)
    obs, _ = env.reset(seed=4)
    p.reset()
    a = p.step(obs["0"])
    nodes_before = p._nodes_bound
    obs, *_ = env.step({"0": a, "1": 0})
    b, lib = p._b, _lib.lib()
    key = np.ascontiguousarray(b.pack_obs(obs["0"]), dtype=np.uint64)
    out4 = np.zeros(4, np.int32)
    lib.posg_tree_advance(b._handle, _lib.ptr(p._arena), p._cap_nodes, p._cap_log, p._root, int(a), _lib.ptr(key), 0,
                          _lib.ptr(out4), None)
    child = int(out4[0])
    assert child > 0, "the real observation was simulated at least once in 6000 simulations"
    before = p._stats(child)
    buf = torch.empty((1 << 16, 2), dtype=torch.int32, device=b.device)
    cnt = np.zeros(1, np.int64)
    lib.posg_tree_gather_belief(b._handle, _lib.ptr(p._arena), p._cap_nodes, p._cap_log, child, _lib.ptr(buf),
                                buf.shape[0], _lib.ptr(cnt), None)
    logged = set(map(tuple, buf[: int(cnt[0])].cpu().numpy().tolist()))
    p.update(a, obs["0"])
    assert p._root == 0 and p.tree_compactions == 1
    after = p._stats(0)
    for k in ("visits", "done", "total"):
        assert (after[k] == before[k]).all(), k
    assert after["node_visits"] == before["node_visits"]
    assert 1 <= p._nodes_bound < nodes_before                      # siblings' subtrees were dropped
    belief = set(map(tuple, p._root_particles.cpu().numpy().tolist()))
    assert logged <= belief and p.root.belief.size() >= p.config.num_particles + p.config.extra_particles
    # searching on from the kept subtree only adds to its statistics
    p.get_action()
    later = p._stats(0)
    assert (later["done"] >= before["done"]).all() and later["done"].sum() == before["done"].sum() + 6000
    p.close()


@pytest.mark.parametrize("wave", [1, 32])
def test_pucb_and_uniform_selection_match_reference(model, wave):
    """action_selection = pucb (PUCB search, max-visit final choice) and uniform (min-visit
    search): root values and visit shares at equal simulation counts vs the reference planner
    (golden: tests/golden/make_planner_modes_golden.py), within 4 SE + 0.02 / + 0.03.  Wave 32
    exercises the closed-form group selection."""
    with open(os.path.join(GOLDEN, "planner_modes_driving.json")) as f:
        modes = json.load(f)["modes"]
    for scen in modes:
        reps = 50
        vals, vis = [], []
        for rep in range(reps):
            p = make_planner(model, seed=rep, max_sims=scen["n_sims"], wave_size=wave, action_selection=scen["mode"],
                             c=scen["c"], pucb_exploration_fraction=0.5)
            p.reset()
            p.update(None, fields_to_obs(scen["ego_obs"]))
            a = p.get_action()
            st = p._last_root_stats
            vals.append(np.where(st["done"] > 0, st["total"] / np.maximum(st["done"], 1), 0.0))
            vis.append(st["done"] / st["done"].sum())
            if scen["mode"] == "pucb":      # final selection = most visited (mcts.py:66-68,566-581)
                assert st["done"][a] == st["done"].max()
            p.close()
        vals, vis = np.array(vals), np.array(vis)
        se_v = np.sqrt(np.array(scen["value_std"]) ** 2 / scen["reps"] + vals.std(0) ** 2 / reps)
        se_n = np.sqrt(np.array(scen["visit_frac_std"]) ** 2 / scen["reps"] + vis.std(0) ** 2 / reps)
        assert (np.abs(vals.mean(0) - np.array(scen["value_mean"])) < 4 * se_v + 0.02).all(), (scen["mode"], vals.mean(0))
        assert (np.abs(vis.mean(0) - np.array(scen["visit_frac_mean"])) < 4 * se_n + 0.03).all(), (scen["mode"], vis.mean(0))
        if scen["mode"] == "uniform" and wave == 1:   # min-visit search: counts differ by at most one
            assert (np.abs(vis * scen["n_sims"] - scen["n_sims"] / 5) <= 1.0 + 1e-9).all()


# ---- PredatorPrey-v0 / LevelBasedForaging-v3 against the unmodified reference planner ----------
@pytest.fixture(scope="module")
def gold_envs():
    with open(os.path.join(GOLDEN, "planner_envs.json")) as f:
        return json.load(f)


def _env_model(env_id, kwargs):
    from posggym_baselines_b200.envs.env import make

    return make(env_id, **kwargs).model


@pytest.mark.parametrize("env_id", ["LevelBasedForaging-v3", "PredatorPrey-v0"])
@pytest.mark.parametrize("wave", [1, 32])
def test_root_statistics_match_reference_other_envs(gold_envs, env_id, wave):
    """As test_root_statistics_match_reference, for the models of BASELINE.json configs[2] and
    configs[4]: root action values after 192 simulations from a first observation within 4
    standard errors + 0.003 of the unmodified reference planner's (values are 0.002 .. 0.05 here),
    visit shares within 0.03, and the reference's clearly best action (LBF, env seed 9: NORTH,
    towards the food in sight) is the CUDA planner's best action."""
    from posggym_baselines_b200.planning import POMCP, MCTSConfig, RandomSearchPolicy

    g = gold_envs[env_id]
    model = _env_model(env_id, gold_envs["env_kwargs"][env_id])
    for scen in g["root_stats"]:
        reps, vals, vis = 60, [], []
        for rep in range(reps):
            cfg = MCTSConfig(discount=0.95, search_time_limit=0.1, c=1.414, truncated=False, action_selection="ucb",
                             seed=rep)
            p = POMCP(model, "0", cfg, RandomSearchPolicy(model, "0"), max_sims=scen["n_sims"], wave_size=wave,
                      node_capacity=1 << 18, log_capacity=1 << 21)
            p.reset()
            p.update(None, tuple(scen["ego_obs"]))
            p.get_action()
            st = p._last_root_stats
            vals.append(np.where(st["done"] > 0, st["total"] / np.maximum(st["done"], 1), 0.0))
            vis.append(st["done"] / st["done"].sum())
            p.close()
        vals, vis = np.array(vals), np.array(vis)
        ref_mean, ref_std = np.array(scen["value_mean"]), np.array(scen["value_std"])
        se = np.sqrt(ref_std**2 / scen["reps"] + vals.std(0) ** 2 / reps)
        assert (np.abs(vals.mean(0) - ref_mean) < 4 * se + 0.003).all(), (env_id, vals.mean(0), ref_mean, se)
        assert (np.abs(vis.mean(0) - np.array(scen["visit_frac_mean"])) < 0.03).all(), (vis.mean(0), scen["visit_frac_mean"])
        order = np.argsort(ref_mean)
        if ref_mean[order[-1]] - ref_mean[order[-2]] > 4 * se.max() + 0.003:
            assert int(np.argmax(vals.mean(0))) == int(order[-1])


@pytest.mark.parametrize("env_id", ["LevelBasedForaging-v3", "PredatorPrey-v0"])
def test_rollout_return_distribution_ks_other_envs(gold_envs, env_id):
    """KS two-sample test of the fused rollout kernel's returns against the reference's
    MCTS._rollout (its own RNG streams) from the same state: not rejected at p > 0.001."""
    from tests.test_lbf_gpu import to_states as lbf_states
    from tests.test_pp_gpu import to_states as pp_states

    scen = gold_envs[env_id]["rollouts"]
    b = _env_model(env_id, gold_envs["env_kwargs"][env_id]).batched
    f = np.array([scen["state"]])
    st = lbf_states(b, f) if env_id.startswith("Level") else pp_states(f)
    packed = torch.from_numpy(np.repeat(b.pack_state(st), 100_000, axis=0)).to(b.device)
    max_steps = min(scen["depth_limit"] + 1, scen["step_limit"] - 1 + 1)
    ret, _, _ = b.rollout(packed, 0, max_steps, scen["discount"], seed=321)
    d, p = ks_2samp(ret.cpu().numpy(), np.array(scen["returns_sorted"]))
    assert p > 0.001, (env_id, d, p)


# ---- the throughput regime: the wave schedule planner.step runs under a time limit ------------------------
@pytest.fixture(scope="module")
def gold_large():
    with open(os.path.join(GOLDEN, "planner_large.json")) as f:
        return json.load(f)


def _first_obs(scen, gold, gold_envs):
    if scen["env_id"] == "Driving-v1":
        return fields_to_obs(next(r for r in gold["root_stats"] if r["env_seed"] == scen["env_seed"])["ego_obs"])
    return tuple(next(r for r in gold_envs[scen["env_id"]]["root_stats"] if r["env_seed"] == scen["env_seed"])["ego_obs"])


@pytest.mark.parametrize("idx", [0, 1, 2, 3])
def test_wave_schedule_root_statistics_at_65536_sims(gold, gold_envs, gold_large, idx):
    """The regime the bench's simulation counts come from: NO fixed wave size -- the planner's own schedule,
    wave = root visits / 2048 (at least one warp), consecutive waves chained inside one launch -- at a LARGE
    equal budget (65 536 simulations from a fixed first observation) against the unmodified reference
    planner's sequential search of the same budget
    (tests/golden/make_planner_large_golden.py; Driving-v1 x2, PredatorPrey-v0, LevelBasedForaging-v3).
    Root action values within 4 standard errors + 0.02 (Driving; + 0.003 for the other two, whose values are
    ~0.003), the visit share of the reference's most visited action within 0.15, the reference's clearly
    best action is the CUDA planner's best action, and no simulation is lost to a full arena."""
    from posggym_baselines_b200.planning import POMCP, MCTSConfig, RandomSearchPolicy

    scen = gold_large["scenarios"][idx]
    env_id = scen["env_id"]
    model = _env_model(env_id, gold_large["env_kwargs"][env_id])
    obs = _first_obs(scen, gold, gold_envs)
    reps, vals, vis = 8, [], []
    for rep in range(reps):
        cfg = MCTSConfig(discount=0.95, search_time_limit=0.1, c=1.414, truncated=False, action_selection="ucb", seed=rep)
        p = POMCP(model, "0", cfg, RandomSearchPolicy(model, "0"), max_sims=gold_large["n_sims"],
                  node_capacity=1 << 18, log_capacity=1 << 22)
        p.reset()
        p.update(None, obs)
        p.get_action()
        st, info = p._last_root_stats, p.search_info
        assert info["num_sims"] == gold_large["n_sims"] and info["last_wave"] <= max(64, gold_large["n_sims"] // 2048)
        assert info["overflow_sims"] == 0 and info["ids_reserved"] == gold_large["n_sims"]
        assert 0 < info["nodes_created"] <= gold_large["n_sims"]
        vals.append(np.where(st["done"] > 0, st["total"] / np.maximum(st["done"], 1), 0.0))
        vis.append(st["done"] / st["done"].sum())
        p.close()
    vals, vis = np.array(vals), np.array(vis)
    ref_mean, ref_std = np.array(scen["value_mean"]), np.array(scen["value_std"])
    tol = 0.02 if env_id == "Driving-v1" else 0.003
    se = np.sqrt(ref_std**2 / gold_large["reps"] + vals.std(0) ** 2 / reps)
    assert (np.abs(vals.mean(0) - ref_mean) < 4 * se + tol).all(), (env_id, vals.mean(0), ref_mean, se)
    ref_vis = np.array(scen["visit_frac_mean"])
    top = int(np.argmax(ref_vis))
    assert abs(vis.mean(0)[top] - ref_vis[top]) < 0.15, (vis.mean(0), ref_vis)
    order = np.argsort(ref_mean)
    if ref_mean[order[-1]] - ref_mean[order[-2]] > 4 * se.max() + tol:
        assert int(np.argmax(vals.mean(0))) == int(order[-1])


def test_full_arena_is_reported_not_hidden(model):
    """A search that outgrows its arena keeps evaluating unexpanded histories by rollouts (search.cuh) --
    and says so: search_info counts the simulations that could not add a node."""
    p = make_planner(model, seed=2, max_sims=40_000, node_capacity=1 << 14, log_capacity=1 << 18)
    p.reset()
    from posggym_baselines_b200.envs.env import Env

    obs, _ = Env(model).reset(seed=2)
    p.update(None, obs["0"])
    p.get_action()
    info = p.search_info
    assert info["num_sims"] == 40_000 and info["overflow_sims"] > 0
    assert info["expanding_sims"] + info["overflow_sims"] == info["num_sims"]
    assert info["arena_fill"] > 0.9
    p.close()


def test_throughput_mode_episodes_vs_sequential_mode(model):
    """Episode outcomes of the throughput mode (time-limited, the planner's own wave schedule: ~1e4 simulations
    per 0.02 s decision) against the sequential mode (wave size 1) at the reference's ~100 simulations per decision:
    more search must not play worse -- success rate not lower by more than 3 binomial standard errors + 0.05
    -- and the return distributions are not told apart by a KS two-sample test at p > 0.001."""
    from posggym_baselines_b200.envs.env import Env

    env = Env(model)

    def play(n, first=0, **opts):
        wins, rets, sims = 0, [], []
        for ep in range(first, first + n):
            planner = make_planner(model, seed=ep, **opts)
            rng = random.Random(20_000 + ep)
            obs, _ = env.reset(seed=1000 + ep)
            planner.reset()
            done, R = False, 0.0
            while not done:
                a = planner.step(obs["0"])
                if planner.step_statistics["num_sims"] == planner.step_statistics["num_sims"]:
                    sims.append(planner.step_statistics["num_sims"])
                obs, rew, term, trunc, done, _ = env.step({"0": a, "1": rng.randrange(5)})
                R += rew["0"]
            wins += env.state[0].dest_reached
            rets.append(R)
            planner.close()
        return wins, np.array(rets), float(np.mean(sims))

    n = 30
    w_seq, r_seq, _ = play(n, max_sims=100, wave_size=1)
    w_thr, r_thr, sims = play(n, search_time_limit=0.02, node_capacity=1 << 24, log_capacity=1 << 26)
    n_thr = n

    # success rates with one pseudo-success and one pseudo-failure each (30 / 30 would otherwise claim zero variance
    # and turn the bound into "never lose more than four episodes")
    p_seq, p_first = (w_seq + 1) / (n + 2), (w_thr + 1) / (n + 2)
    limit = p_seq - 3 * math.sqrt(p_seq * (1 - p_seq) / n + p_first * (1 - p_first) / n) - 0.05
    if p_first <= limit:
        # the throughput mode is WALL-CLOCK limited (0.02 s per decision): a box that stalls for a moment hands some
        # decisions a fraction of their simulations.  One more batch of episodes before calling it a difference.
        w2, r2, sims2 = play(n, first=n, search_time_limit=0.02, node_capacity=1 << 24, log_capacity=1 << 26)
        w_thr, r_thr, n_thr, sims = w_thr + w2, np.concatenate([r_thr, r2]), 2 * n, (sims + sims2) / 2
    assert (w_thr + 1) / (n_thr + 2) > limit, (w_thr / n_thr, w_seq / n, limit, sims)
    d, pv = ks_2samp(r_thr, r_seq)
    assert pv > 0.001, (d, pv, r_thr.mean(), r_seq.mean(), sims)


@pytest.mark.parametrize("growth,ok", [(1.0 / 2048, True), (1.0, False)])
def test_wave_size_matters_at_large_budgets(gold, gold_large, growth, ok):
    """Why the schedule is what it is: on the first Driving scenario the reference's search locks on to
    ACCELERATE between 40 000 and 65 536 simulations (value 0.06, 56 % of the visits).  Waves of 1/2048 of
    the root's visits reproduce that; waves that double the tree (wave = N, 12 launches for the whole budget
    -- the schedule round 1 benchmarked) leave all five actions near the random-rollout value."""
    from posggym_baselines_b200.planning import POMCP, MCTSConfig, RandomSearchPolicy

    scen = gold_large["scenarios"][0]
    model = _env_model("Driving-v1", gold_large["env_kwargs"]["Driving-v1"])
    obs = _first_obs(scen, gold, None)
    vals = []
    for rep in range(4):
        cfg = MCTSConfig(discount=0.95, search_time_limit=0.1, c=1.414, truncated=False, action_selection="ucb", seed=rep)
        p = POMCP(model, "0", cfg, RandomSearchPolicy(model, "0"), max_sims=gold_large["n_sims"], wave_growth=growth,
                  node_capacity=1 << 18, log_capacity=1 << 22)
        p.reset()
        p.update(None, obs)
        p.get_action()
        st = p._last_root_stats
        vals.append(np.where(st["done"] > 0, st["total"] / np.maximum(st["done"], 1), 0.0))
        p.close()
    gap = np.mean(vals, 0)[1] - np.delete(np.mean(vals, 0), 1).max()
    ref_gap = scen["value_mean"][1] - np.delete(np.array(scen["value_mean"]), 1).max()
    assert ref_gap > 0.08
    assert (gap > 0.5 * ref_gap) == ok, (growth, gap, ref_gap)
