"""The GPU vector env against the reference's vec-env stack restated on the CPU oracle
(oracle/vec_env.py): same shapes / dtypes / auto-reset / episode statistics, and -- with
the model noise replayed -- the same numbers."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_contract_shapes_and_autoreset():
    from posggym_baselines_b200.vector import load_vec_env

    n = 4096
    envs = load_vec_env("Driving-v1", {"grid": "14x14RoundAbout", "num_agents": 2, "obs_dim": (3, 1, 1)}, num_envs=n, seed=1)
    obs, infos = envs.reset()
    assert obs.shape == (n * 2, 124) and obs.dtype == torch.float32
    assert set(infos) == set(envs.possible_agents) == {"0", "1"}
    # every one-hot block sums to one: 15 cells + speed + 4 coords + 2 flags = 22 ones per row
    assert torch.equal(obs.sum(1), torch.full((n * 2,), 22.0, device=obs.device))
    g = torch.Generator(device="cpu").manual_seed(0)
    ep_lens, ep_rets, steps = [], [], 0
    for _ in range(120):
        act = torch.randint(0, 5, (n * 2,), generator=g)
        obs, rew, term, trunc, dones, infos = envs.step(act.numpy())
        steps += 1
        assert obs.shape == (n * 2, 124) and rew.shape == (n * 2,) and rew.dtype == torch.float32
        assert term.shape == trunc.shape == (n * 2,) and dones.shape == (n,)
        assert torch.equal(obs.sum(1), torch.full((n * 2,), 22.0, device=obs.device))
        if dones.any():
            ep = infos["0"]["episode"]
            assert ep["r"].shape == (n,) and ep["l"].shape == (n,)
            ep_lens.append(ep["l"][dones].cpu())
            ep_rets.append(ep["r"][dones].cpu())
            # truncation fires exactly at the 50-step limit, for both agents of the env
            tr = trunc.reshape(n, 2)
            assert ((ep["l"] == 50) == tr[:, 0])[dones].all()
            # auto-reset: the returned observation is a first observation (speed STOPPED, not done)
            first = obs.reshape(n, 2, 124)[dones]
            assert (first[:, :, 61] == 1).all() and (first[:, :, 120] == 1).all() and (first[:, :, 122] == 1).all()
    lens = torch.cat(ep_lens)
    assert lens.max() == 50 and lens.min() >= 1 and (lens == 50).float().mean() > 0.6
    rets = torch.cat(ep_rets)
    # uniform-random joint play (div_results.csv): mean return about -0.1, crashes in ~20 % of episodes
    assert -0.25 < rets.mean().item() < 0.05


def test_matches_cpu_vec_env_stack(oracle_env):
    """Replay: the GPU env and the oracle stack are driven with the same actions and the
    same model noise / initial states; observations, rewards, flags and episode statistics
    must be identical at every step (float32 rewards as StackEnv returns them)."""
    import posggym
    from gymnasium import spaces

    from oracle import rollout as R
    from oracle.driving import DrivingModel
    from oracle.vec_env import FlattenObservations
    from posggym_baselines_b200.vector import load_vec_env

    n, seed = 96, 5
    envs = load_vec_env("Driving-v1", num_envs=n, seed=seed)
    obs, _ = envs.reset()
    m = DrivingModel()
    sp = m.observation_spaces["0"]

    def first_state(env, ep):
        s = seed ^ ((0x9E3779B97F4A7C15 * ep) & 0xFFFFFFFFFFFFFFFF) if ep else seed
        return m.initial_state_from_noise(R.driving_initial_picks(m, s, np.array([env], dtype=np.uint64))[0])

    states = [first_state(e, 0) for e in range(n)]
    eps = [0] * n
    t = [0] * n
    ret = np.zeros((n, 2))
    flat = lambda o: np.concatenate([spaces.flatten(sp, o["0"]), spaces.flatten(sp, o["1"])]).reshape(2, -1)  # noqa: E731
    cpu_obs = np.stack([flat(m.sample_initial_obs(s)) for s in states]).reshape(n * 2, -1)
    assert (obs.cpu().numpy() == cpu_obs).all()
    rng = np.random.default_rng(0)
    for step in range(1, 131):
        act = rng.integers(0, 5, size=(n, 2))
        obs, rew, term, trunc, dones, infos = envs.step(act.reshape(-1))
        _, order = R.driving_step_noise(seed, np.arange(n, dtype=np.uint64), step)
        c_obs, c_rew, c_term, c_trunc, c_done = [], [], [], [], []
        c_r, c_l = np.zeros((n, 2), np.float32), np.zeros(n, np.int32)
        for e in range(n):
            ts = m.step_with_noise(states[e], {"0": int(act[e, 0]), "1": int(act[e, 1])}, int(order[e]))
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


def test_predator_prey_vec_env_and_sharding():
    from posggym_baselines_b200.vector import load_vec_env

    n = 2048
    full = load_vec_env("PredatorPrey-v0", num_envs=n, seed=3)
    lo = load_vec_env("PredatorPrey-v0", num_envs=n // 2, seed=3, env_offset=0)
    hi = load_vec_env("PredatorPrey-v0", num_envs=n // 2, seed=3, env_offset=n // 2)
    o, ol, oh = full.reset()[0], lo.reset()[0], hi.reset()[0]
    assert o.shape == (n * 2, 100)
    assert torch.equal(o, torch.cat([ol, oh]))
    g = torch.Generator().manual_seed(1)
    for _ in range(60):
        act = torch.randint(0, 5, (n * 2,), generator=g)
        a = full.step(act)
        b = lo.step(act[: n])
        c = hi.step(act[n:])
        for k in range(5):
            assert torch.equal(a[k], torch.cat([b[k], c[k]])), k


@pytest.mark.parametrize("env_id,nact", [("LevelBasedForaging-v3", 6), ("Driving-v1", 5), ("PredatorPrey-v0", 5)])
def test_embed_matches_torch_linear_on_one_hot_rows(env_id, nact):
    """posg_vec_embed == tanh(F.linear(one_hot_obs, W, b)) of the first trunk layer
    (ppo/network.py:171-178) in float32: the one-hot product is a sum of one weight column per
    observation field.  Tolerance 2e-6 absolute (float32 sums of <= 31 terms in a different order),
    written here because this is the one floating-point kernel of the path.  Also: an env built
    with write_obs=False steps identically and never writes the one-hot rows."""
    from posggym_baselines_b200.vector import load_vec_env

    n, hidden = 3000, 64
    envs = load_vec_env(env_id, num_envs=n, seed=11)
    lean = load_vec_env(env_id, num_envs=n, seed=11)
    lean.write_obs, lean._obs, lean._args = False, None, None
    g = torch.Generator().manual_seed(3)
    w = (torch.randn(hidden, envs.flat_dim, generator=g) * 0.3).cuda()
    b = (torch.randn(hidden, generator=g) * 0.1).cuda()
    obs = envs.reset()[0]
    assert lean.reset()[0] is None
    for step in range(12):
        ref = torch.tanh(torch.nn.functional.linear(obs, w, b))
        got, got_lean = envs.embed(w, b), lean.embed(w, b)
        assert got.shape == (2 * n, hidden) and got.dtype == torch.float32
        assert (got - ref).abs().max().item() < 2e-6, (step, (got - ref).abs().max().item())
        assert torch.equal(got, got_lean)
        raw = envs.embed(w, b, tanh=False)
        assert (raw - torch.nn.functional.linear(obs, w, b)).abs().max().item() < 5e-6
        act = torch.randint(0, nact, (2 * n,), generator=g)
        obs, rew, *_ = envs.step(act)
        o2, rew2, *_ = lean.step(act)
        assert o2 is None and torch.equal(rew, rew2)
    # a hidden size that is not a multiple of 32
    w2, b2 = w[:40].contiguous(), b[:40].contiguous()
    assert (envs.embed(w2, b2) - torch.tanh(torch.nn.functional.linear(obs, w2, b2))).abs().max().item() < 2e-6


@pytest.mark.parametrize("env_id,nact", [("Driving-v1", 5), ("LevelBasedForaging-v3", 6)])
def test_captured_step_equals_eager_step(env_id, nact):
    """envs.step captured in a CUDA graph (step index read from device memory) == the eager call,
    also when eager steps are mixed in; Driving's processing-order noise depends on the step index."""
    from posggym_baselines_b200.vector import load_vec_env

    n = 4096
    eager = load_vec_env(env_id, num_envs=n, seed=21)
    graphed = load_vec_env(env_id, num_envs=n, seed=21)
    assert torch.equal(eager.reset()[0], graphed.reset()[0])
    actions, replay = graphed.capture_step()
    g = torch.Generator().manual_seed(5)
    for step in range(80):
        act = torch.randint(0, nact, (2 * n,), generator=g).cuda()
        a = eager.step(act)
        if step in (20, 21, 50):       # an eager step in between must not desynchronise the counter
            b = graphed.step(act)
        else:
            actions.copy_(act)
            b = replay()
        for k in range(5):
            assert torch.equal(a[k], b[k]), (step, k)
        assert torch.equal(a[5]["0"]["episode"]["r"], b[5]["0"]["episode"]["r"])
        assert torch.equal(a[5]["1"]["episode"]["l"], b[5]["1"]["episode"]["l"])
    # a whole rollout in one graph (step_device), replayed twice, against eager steps
    acts = torch.randint(0, nact, (16, 2 * n), generator=g).cuda()
    roll = torch.cuda.CUDAGraph()
    with torch.cuda.graph(roll):
        for k in range(16):
            out = graphed.step_device(acts[k])
    for rep in range(2):
        roll.replay()
        for k in range(16):
            ref = eager.step(acts[k])
        for k in range(5):
            assert torch.equal(ref[k], out[k]), (rep, k)
    # reset() restarts the device counter as well
    assert torch.equal(eager.reset()[0], graphed.reset()[0])
    a, b = eager.step(acts[0]), graphed.step(acts[0])
    for k in range(5):
        assert torch.equal(a[k], b[k])


@pytest.mark.parametrize("env_id,nact", [("LevelBasedForaging-v3", 6), ("Driving-v1", 5), ("PredatorPrey-v0", 5)])
def test_fused_step_embed_equals_step_then_embed(env_id, nact):
    """posg_vec_step_embed (the collector's second launch) == envs.step followed by posg_vec_embed, bit for bit:
    rewards, terminated / truncated, the carried state, the copy of the state for the rollout storage and the
    embedding of the observations the step returned -- over enough steps for episodes to end and auto-reset."""
    from posggym_baselines_b200.vector import load_vec_env

    n, hidden = 2048 + 37, 64
    a = load_vec_env(env_id, num_envs=n, seed=13)
    b = load_vec_env(env_id, num_envs=n, seed=13)
    for e in (a, b):
        e.write_obs, e._obs, e._args = False, None, None
        e.reset()
        e.enable_device_counter()
    g = torch.Generator().manual_seed(4)
    w = (torch.randn(hidden, a.flat_dim, generator=g) * 0.3).cuda()
    w_t, bias = w.t().contiguous(), (torch.randn(hidden, generator=g) * 0.1).cuda()
    x = torch.empty((2 * n, hidden), dtype=torch.float32, device="cuda")
    rew, term, trunc = (torch.empty(2 * n, dtype=dt, device="cuda") for dt in (torch.float32, torch.uint8, torch.uint8))
    copy = torch.empty_like(b._state)
    for step in range(70):
        act = torch.randint(0, nact, (2 * n,), generator=g).cuda()
        a.step_device(act)
        ref_x = a.embed(w, bias)
        b.step_embed_device(act, w_t, bias, x, outs=(rew, term, trunc), state_copy=copy)
        assert torch.equal(a._state, b._state) and torch.equal(copy, b._state), step
        assert torch.equal(a._rew, rew) and torch.equal(a._term, term) and torch.equal(a._trunc, trunc), step
        assert torch.equal(a._done, b._done) and torch.equal(a._stat_r, b._stat_r) and torch.equal(a._stat_l, b._stat_l)
        assert torch.equal(ref_x, x), (step, (ref_x - x).abs().max().item())
    assert int(a._ep_count.sum()) > 0      # episodes ended and were reset on the way
