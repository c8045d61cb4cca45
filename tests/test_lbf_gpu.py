"""Parity of the LevelBasedForaging-v3 CUDA path (through the C-ABI) against the committed
golden vectors and the CPU oracle: integer state / observations / flags and float64 rewards
/ returns BIT-EXACT; the PPO vec-env (BASELINE.json configs[4]) replayed against the
reference's vec-env stack restated on the oracle; POMCP smoke on the model."""
import random

import numpy as np
import pytest
import torch

from tests.conftest import golden

pytestmark = pytest.mark.gpu

KW = dict(num_agents=2, max_agent_level=3, size=10, max_food=8, sight=2, force_coop=False, static_layout=True,
          observation_mode="tuple")   # baseline_exps/env_data/LevelBasedForaging-v3/env_kwargs.yaml


@pytest.fixture(scope="module")
def env():
    from posggym_baselines_b200.envs.lbf import BatchedLBF

    return BatchedLBF(**KW)


def to_states(env, fields):
    from posggym_baselines_b200.envs.lbf import Food, LBFState, Player

    out = []
    for f in fields.tolist():
        out.append(LBFState((Player(0, (f[0], f[1]), f[2]), Player(1, (f[3], f[4]), f[5])),
                            tuple(Food(env.food_coords[k], f[6 + k]) for k in range(8)), f[14]))
    return out


def fields(states):
    return np.array([[*[v for p in s.players for v in (p.coord[0], p.coord[1], p.level)], *[f.level for f in s.food],
                      s.food_spawned] for s in states], np.int8)


def test_step_matches_golden(env):
    g = golden("lbf_step.npz")
    dev = env.device
    packed = torch.from_numpy(env.pack_state(to_states(env, g["state"]))).to(dev)
    noise = torch.zeros((len(g["state"]), 1), dtype=torch.int32, device=dev)
    out = env.step(packed, torch.from_numpy(g["action"]).to(dev), noise)
    torch.cuda.synchronize()
    got = fields(env.unpack_state(out.state))
    assert (got == g["next_state"]).all(), np.nonzero((got != g["next_state"]).any(1))[0][:10]
    obs = np.array(env.unpack_obs(out.obs), np.int8)
    assert (obs == g["obs"]).all(), np.nonzero((obs != g["obs"]).any((1, 2)))[0][:10]
    assert (out.rewards.cpu().numpy().view(np.uint64) == g["reward"].view(np.uint64)).all()
    flags = out.flags.cpu().numpy()
    assert ((flags & 0x80).astype(bool) == g["all_done"]).all()
    assert (((flags & 3) == 3) == g["all_done"]).all()
    # the set exercises single loads, joint loads, failed loads and the terminal transition
    assert (g["reward"] > 0).all(1).any() and g["all_done"].any()


def test_obs_codec_round_trip(env):
    g = golden("lbf_step.npz")
    for o in g["obs"][:500].reshape(-1, 30):
        w = env.pack_obs([int(v) for v in o])
        assert env.unpack_obs(w.view(np.int64).reshape(1, 1, 2))[0][0] == tuple(int(v) for v in o)


def test_initial_state_and_obs_match_golden(env):
    g = golden("lbf_init.npz")
    st = env.sample_initial_state(len(g["picks"]), picks=torch.from_numpy(g["picks"]).to(env.device))
    assert (fields(env.unpack_state(st)) == g["state"]).all()
    assert (np.array(env.unpack_obs(env.observe(st)), np.int8) == g["obs"]).all()


def test_rollout_matches_golden(env):
    g = golden("lbf_rollout.npz")
    seed, disc, max_steps, ego = int(g["seed"]), float(g["discount"]), int(g["max_steps"]), int(g["ego"])
    n, off = len(g["particle"]), int(g["particle"][0])
    st = env.sample_initial_state(n, seed=seed, particle_offset=off)
    assert (fields(env.unpack_state(st)) == g["init_state"]).all()
    ret, steps, final = env.rollout(st, ego, max_steps, disc, seed=seed, particle_offset=off, want_state=True)
    assert (ret.cpu().numpy().view(np.uint64) == g["ret"].view(np.uint64)).all()
    assert (steps.cpu().numpy() == g["steps"]).all()
    assert (fields(env.unpack_state(final)) == g["final_state"]).all()


def test_step_vs_oracle_live(env, oracle_env):
    from oracle import rollout as R
    from oracle.lbf import LevelBasedForagingModel

    m = LevelBasedForagingModel(**KW)
    n, seed, off = 3000, 77, 12345
    parts = np.arange(n, dtype=np.uint64) + np.uint64(off)
    st = env.sample_initial_state(n, seed=seed, particle_offset=off)
    cpu = [m.initial_state_from_noise([int(v) for v in pk]) for pk in R.lbf_initial_picks(m, seed, parts)]
    assert fields(env.unpack_state(st)).tolist() == fields(cpu).tolist()
    cur = st
    for step in range(25):
        out = env.step(cur, seed=seed, step_idx=step, particle_offset=off)
        acts = R.lbf_step_noise(seed, parts, step)
        assert (out.actions.cpu().numpy() == acts).all()
        gs, go = fields(env.unpack_state(out.state)), env.unpack_obs(out.obs)
        rew, flags = out.rewards.cpu().numpy(), out.flags.cpu().numpy()
        for i in range(n):
            ts = m.step(cpu[i], {"0": int(acts[i, 0]), "1": int(acts[i, 1])})
            cpu[i] = ts.state
            assert go[i] == (ts.observations["0"], ts.observations["1"]), (step, i)
            assert rew[i, 0] == ts.rewards["0"] and rew[i, 1] == ts.rewards["1"]
            assert bool(flags[i] & 0x80) == ts.all_done
        assert gs.tolist() == fields(cpu).tolist(), step
        cur = out.state


def test_fast_path_vs_oracle_and_general_kernel(env, oracle_env):
    """The random-action sweep runs whole 128-env blocks through generic::step_fast_kernel (offset a
    multiple of 4) and the ragged tail through the general kernel: every output against the oracle
    and against the general kernel driven with the same explicit actions."""
    from oracle import rollout as R
    from oracle.lbf import LevelBasedForagingModel

    m = LevelBasedForagingModel(**KW)
    n, seed, off = 1000, 6, 4 * 321          # 7 full blocks + a 104-env tail
    parts = np.arange(n, dtype=np.uint64) + np.uint64(off)
    st = env.sample_initial_state(n, seed=seed, particle_offset=off)
    cpu = [m.initial_state_from_noise([int(v) for v in pk]) for pk in R.lbf_initial_picks(m, seed, parts)]
    live = torch.zeros(1, dtype=torch.int64, device=env.device)
    cur, expect_live = st, 0
    zero_noise = torch.zeros((n, 1), dtype=torch.int32, device=env.device)
    for step in range(30):
        out = env.step(cur, seed=seed, step_idx=step, particle_offset=off, live_count=live)
        acts = R.lbf_step_noise(seed, parts, step)
        assert (out.actions.cpu().numpy() == acts).all()
        ref = env.step(cur, torch.from_numpy(acts.astype(np.uint8)).to(env.device), zero_noise)
        assert torch.equal(out.state, ref.state) and torch.equal(out.obs, ref.obs) and torch.equal(out.flags, ref.flags)
        assert torch.equal(out.rewards.view(torch.int64), ref.rewards.view(torch.int64))
        lean = env.step(cur, seed=seed, step_idx=step, particle_offset=off, want_actions=False)
        assert torch.equal(out.state, lean.state) and torch.equal(out.obs, lean.obs)
        go, rew, flags = env.unpack_obs(out.obs), out.rewards.cpu().numpy(), out.flags.cpu().numpy()
        for i in range(n):
            expect_live += any(f.level > 0 for f in cpu[i].food)
            ts = m.step(cpu[i], {"0": int(acts[i, 0]), "1": int(acts[i, 1])})
            cpu[i] = ts.state
            assert go[i] == (ts.observations["0"], ts.observations["1"]), (step, i)
            assert rew[i, 0] == ts.rewards["0"] and rew[i, 1] == ts.rewards["1"]
            assert bool(flags[i] & 0x80) == ts.all_done
        assert fields(env.unpack_state(out.state)).tolist() == fields(cpu).tolist(), step
        cur = out.state
    assert int(live.item()) == expect_live


def test_sweep_equals_rollout_at_scale(env):
    """Full-size property: the fused rollout equals a 50-launch step sweep, and a shard by
    particle offset equals the monolithic run (2 M particles x 50 steps)."""
    n, depth, seed, disc = 1 << 21, 50, 505, 0.95
    st = env.sample_initial_state(n, seed=seed)
    ret, steps, final = env.rollout(st, 0, depth, disc, seed=seed, want_state=True)
    cur = st.clone()
    acc = torch.zeros(n, dtype=torch.float64, device=env.device)
    alive = torch.ones(n, dtype=torch.bool, device=env.device)
    frozen = st.clone()
    tot = torch.zeros((n, 2), dtype=torch.float64, device=env.device)
    for k in range(depth):
        out = env.step(cur, seed=seed, step_idx=k, want_actions=False)
        acc = torch.where(alive, acc + (disc**k) * out.rewards[:, 0], acc)
        tot += torch.where(alive.unsqueeze(1), out.rewards, torch.zeros_like(out.rewards))
        frozen = torch.where(alive.unsqueeze(1), out.state, frozen)
        alive = alive & ((out.flags & 0x81) == 0)
        cur = out.state
    assert torch.equal(acc, ret) and torch.equal(frozen, final)
    h = n // 2
    r2, s2, _ = env.rollout(st[h:].contiguous(), 0, depth, disc, seed=seed, particle_offset=h)
    assert torch.equal(r2, ret[h:]) and torch.equal(s2, steps[h:])
    # uniform-random joint play (div_results.csv, 100 episodes): 0.0597 +- 0.0088 / 0.0158 +- 0.0041
    mean = tot.mean(0).tolist()
    assert abs(mean[0] - 0.0597) < 3 * 0.0088 and abs(mean[1] - 0.0158) < 3 * 0.0041, mean
    # every reward is level * food_level / (loaders' level sum * spawned): an episode never pays more than 1
    assert float(tot.sum(1).max()) <= 1.0 + 1e-12


def test_vec_env_matches_cpu_stack(oracle_env):
    """PPO collector contract (ppo/worker.py:82-86,176-216) on LBF: replay against the oracle
    with the same actions; 300-wide one-hot float32 observations, float32 rewards, flags and
    episode statistics identical at every step."""
    from gymnasium import spaces

    from oracle import rollout as R
    from oracle.lbf import LevelBasedForagingModel
    from posggym_baselines_b200.vector import load_vec_env

    n, seed = 64, 9
    envs = load_vec_env("LevelBasedForaging-v3", KW, num_envs=n, seed=seed)
    obs, _ = envs.reset()
    assert obs.shape == (n * 2, 300) and obs.dtype == torch.float32
    m = LevelBasedForagingModel(**KW)
    sp = m.observation_spaces["0"]

    def first_state(env, ep):
        s = seed ^ ((0x9E3779B97F4A7C15 * ep) & 0xFFFFFFFFFFFFFFFF) if ep else seed
        return m.initial_state_from_noise([int(v) for v in R.lbf_initial_picks(m, s, np.array([env], dtype=np.uint64))[0]])

    states = [first_state(e, 0) for e in range(n)]
    eps, t, ret = [0] * n, [0] * n, np.zeros((n, 2))
    flat = lambda o: np.stack([spaces.flatten(sp, o["0"]), spaces.flatten(sp, o["1"])])  # noqa: E731
    cpu_obs = np.stack([flat(m.sample_initial_obs(s)) for s in states]).reshape(n * 2, -1)
    assert (obs.cpu().numpy() == cpu_obs).all()
    rng = random.Random(0)
    from tests.golden.make_golden import lbf_greedy_action

    for step in range(1, 121):
        # half the envs forage greedily so that episodes also end by termination
        act = np.array([[lbf_greedy_action(m, states[e], i, rng) if e % 2 else rng.randrange(6) for i in range(2)]
                        for e in range(n)])
        obs, rew, term, trunc, dones, infos = envs.step(act.reshape(-1))
        c_obs, c_rew, c_term, c_trunc, c_done = [], [], [], [], []
        c_r, c_l = np.zeros((n, 2), np.float32), np.zeros(n, np.int32)
        for e in range(n):
            ts = m.step(states[e], {"0": int(act[e, 0]), "1": int(act[e, 1])})
            t[e] += 1
            tr = t[e] >= 50
            d = ts.all_done or tr
            ret[e] += [ts.rewards["0"], ts.rewards["1"]]
            c_rew.append([np.float32(ts.rewards["0"]), np.float32(ts.rewards["1"])])
            c_term.append([ts.terminations["0"], ts.terminations["1"]])
            c_trunc.append([tr, tr]); c_done.append(d)
            if d:
                c_r[e], c_l[e] = ret[e].astype(np.float32), t[e]
                eps[e] += 1
                states[e] = first_state(e, eps[e]); t[e] = 0; ret[e] = 0
                c_obs.append(flat(m.sample_initial_obs(states[e])))
            else:
                states[e] = ts.state
                c_obs.append(flat(ts.observations))
        assert (obs.cpu().numpy() == np.stack(c_obs).reshape(n * 2, -1)).all(), step
        assert (rew.cpu().numpy() == np.array(c_rew, np.float32).reshape(-1)).all(), step
        assert (term.cpu().numpy() == np.array(c_term).reshape(-1)).all()
        assert (trunc.cpu().numpy() == np.array(c_trunc).reshape(-1)).all()
        assert (dones.cpu().numpy() == np.array(c_done)).all()
        if any(c_done):
            assert (infos["0"]["episode"]["l"].cpu().numpy() == c_l).all()
            assert (infos["0"]["episode"]["r"].cpu().numpy() == c_r[:, 0]).all()
            assert (infos["1"]["episode"]["r"].cpu().numpy() == c_r[:, 1]).all()
    assert any(e > 0 for e in eps)


def test_vec_env_sharding_65536():
    """configs[4]: 65 536 envs sharded in 8 blocks of 8 192 == one 65 536-env job."""
    from posggym_baselines_b200.vector import load_vec_env

    n, shards = 65536, 8
    full = load_vec_env("LevelBasedForaging-v3", KW, num_envs=n, seed=4)
    parts = [load_vec_env("LevelBasedForaging-v3", KW, num_envs=n // shards, seed=4, env_offset=k * (n // shards))
             for k in range(shards)]
    o = full.reset()[0]
    assert o.shape == (n * 2, 300)
    assert torch.equal(o, torch.cat([p.reset()[0] for p in parts]))
    # 30 one-hot blocks per row
    assert torch.equal(o.sum(1), torch.full((n * 2,), 30.0, device=o.device))
    g = torch.Generator().manual_seed(2)
    per = 2 * n // shards
    for _ in range(64):
        act = torch.randint(0, 6, (n * 2,), generator=g)
        a = full.step(act)
        bs = [p.step(act[k * per:(k + 1) * per]) for k, p in enumerate(parts)]
        for k in range(5):
            assert torch.equal(a[k], torch.cat([b[k] for b in bs])), k


def test_pomcp_on_lbf():
    from posggym_baselines_b200.envs.env import make
    from posggym_baselines_b200.planning import POMCP, MCTSConfig, RandomSearchPolicy

    env = make("LevelBasedForaging-v3", **KW)
    cfg = MCTSConfig(discount=0.95, search_time_limit=0.05, c=1.414, truncated=False, action_selection="ucb", seed=0)
    planner = POMCP(env.model, "0", cfg, RandomSearchPolicy(env.model, "0"), node_capacity=1 << 20, log_capacity=1 << 23)
    obs, _ = env.reset(seed=3)
    planner.reset()
    rng = random.Random(0)
    total = 0.0
    for _ in range(12):
        a = planner.step(obs["0"])
        assert a in range(6) and planner.step_statistics["num_sims"] > 1000
        obs, rew, _, _, done, _ = env.step({"0": a, "1": rng.randrange(6)})
        total += rew["0"]
        if done:
            break
    planner.close()


def test_edge_cases(env):
    """Empty batch, a single env, stepping a terminal state (all food loaded) and LOAD next to a
    food that is too heavy: nothing changes, nothing is paid."""
    from posggym_baselines_b200.envs.lbf import Food, LBFState, Player

    dev = env.device
    empty = torch.empty((0, 2), dtype=torch.int32, device=dev)
    out = env.step(empty, seed=1)
    assert out.state.shape == (0, 2) and out.obs.shape == (0, 2, 2)
    ret, steps, _ = env.rollout(empty, 0, 50, 0.95)
    assert ret.shape == (0,)
    one = env.sample_initial_state(1, seed=3)
    assert env.step(one, seed=3).state.shape == (1, 2)
    fc = env.food_coords
    # terminal state: every food collected
    done = LBFState((Player(0, (0, 9), 2), Player(1, (9, 0), 3)), tuple(Food(c, 0) for c in fc), 20)
    # player 0 (level 1) next to food 0 (level 3): a failed load
    heavy = LBFState((Player(0, (fc[0][0] - 1, fc[0][1]), 1), Player(1, (9, 0), 1)),
                     tuple(Food(c, 3 if k == 0 else 0) for k, c in enumerate(fc)), 3)
    # both players next to food 0 (level 3), levels 1 + 2: a joint load that pays 1/3 and 2/3
    joint = LBFState((Player(0, (fc[0][0] - 1, fc[0][1]), 1), Player(1, (fc[0][0], fc[0][1] - 1), 2)),
                     tuple(Food(c, 3 if k == 0 else 0) for k, c in enumerate(fc)), 3)
    packed = torch.from_numpy(env.pack_state([done, heavy, joint])).to(dev)
    acts = torch.tensor([[5, 5], [5, 0], [5, 5]], dtype=torch.uint8, device=dev)
    out = env.step(packed, acts, torch.zeros((3, 1), dtype=torch.int32, device=dev))
    st = env.unpack_state(out.state)
    rew, flags = out.rewards.cpu().numpy(), out.flags.cpu().numpy()
    assert st[0] == done and (rew[0] == 0).all() and flags[0] == 0x83
    assert st[1] == heavy and (rew[1] == 0).all() and flags[1] == 0
    assert all(f.level == 0 for f in st[2].food) and flags[2] == 0x83
    assert rew[2, 0] == 1 * 3 / (3 * 3) and rew[2, 1] == 2 * 3 / (3 * 3)
