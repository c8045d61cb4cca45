"""``posg_model_step_host`` (the scalar / small-batch ``model.step`` of the planners' plugin path, host rows in,
host rows out) against the device-buffer entry points on the same rows, bit for bit, for the three models."""
import time

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _models():
    from posggym_baselines_b200.envs.driving import DrivingModel
    from posggym_baselines_b200.envs.lbf import LBFModel
    from posggym_baselines_b200.envs.predator_prey import PredatorPreyModel

    return {"Driving-v1": DrivingModel, "PredatorPrey-v0": PredatorPreyModel, "LevelBasedForaging-v3": LBFModel}


@pytest.mark.parametrize("env_id", ["Driving-v1", "PredatorPrey-v0", "LevelBasedForaging-v3"])
@pytest.mark.parametrize("n", [1, 37, 5000])
def test_step_host_equals_device_step(env_id, n):
    model = _models()[env_id]()
    b = model.batched
    model.seed(n)
    state = b.sample_initial_state(n, seed=5)
    for _ in range(3):                                   # a few random steps away from the initial states
        state = b.step(state, seed=5, step_idx=_).state
    n_act = model.action_spaces["0"].n
    rng = np.random.default_rng(n)
    act = rng.integers(0, n_act, size=(n, len(model.possible_agents)), dtype=np.uint8)
    noise = model.host_step_noise(n)
    dev = b.step(state, torch.from_numpy(act).to(b.device),
                 torch.from_numpy(noise if noise.dtype == np.uint8 else noise.view(np.int32)).to(b.device))
    st, ob, rw, fl = b.step_host(state.cpu().numpy(), act, noise)
    assert np.array_equal(st.reshape(n, -1), dev.state.cpu().numpy().reshape(n, -1))
    assert np.array_equal(ob.reshape(n, -1), dev.obs.cpu().numpy().reshape(n, -1))
    assert np.array_equal(rw.view(np.uint64), dev.rewards.cpu().numpy().view(np.uint64))       # bit-equal doubles
    assert np.array_equal(fl, dev.flags.cpu().numpy())


@pytest.mark.parametrize("env_id", ["Driving-v1", "PredatorPrey-v0", "LevelBasedForaging-v3"])
def test_scalar_facade_step_matches_oracle_and_is_one_round_trip(oracle_env, env_id):
    """The scalar ``model.step`` (mcts.py:333) through the host entry point: same values and the same ``model.rng``
    draws as the oracle model; its cost is printed (one staged copy in, one kernel, one copy out)."""
    import random

    from oracle.driving import DrivingModel as OracleDriving
    from oracle.lbf import LevelBasedForagingModel as OracleLBF
    from oracle.predator_prey import PredatorPreyModel as OraclePP

    cpu = {"Driving-v1": OracleDriving, "PredatorPrey-v0": OraclePP, "LevelBasedForaging-v3": OracleLBF}[env_id]()
    gpu = _models()[env_id]()
    for m in (cpu, gpu):
        m.seed(3)
    sc, sg = cpu.sample_initial_state(), gpu.sample_initial_state()
    rng = random.Random(0)
    n_act = gpu.action_spaces["0"].n
    for t in range(40):
        a = {i: rng.randrange(n_act) for i in gpu.possible_agents}
        tc, tg = cpu.step(sc, a), gpu.step(sg, a)
        assert tc.rewards == tg.rewards and tc.terminations == tg.terminations and tc.all_done == tg.all_done
        assert _plain(tc.observations) == _plain(tg.observations) and _plain(tc.state) == _plain(tg.state)
        sc, sg = tc.state, tg.state
        if tc.all_done:
            sc, sg = cpu.sample_initial_state(), gpu.sample_initial_state()
    t0 = time.perf_counter()
    for _ in range(300):
        gpu.step(sg, a)
    dt = (time.perf_counter() - t0) / 300
    print(f"{env_id}: scalar model.step {dt * 1e6:.1f} us per call")
    assert dt < 2e-3


def _plain(x):
    if isinstance(x, dict):
        return tuple(_plain(x[k]) for k in sorted(x))
    if isinstance(x, (tuple, list)):
        return tuple(_plain(v) for v in x)
    return int(x)


def test_nested_handle_does_not_share_the_staging():
    """posg_nested_create copies the model handle (I-NTMCP's nested search adapter); the copy must not take the base
    handle's lazily created host staging with it -- closing the planner used to free the block the facade's scalar
    step still used."""
    import math

    from posggym_baselines_b200.planning import INTMCP, MCTSConfig

    model = _models()["Driving-v1"]()
    model.seed(0)
    s = model.sample_initial_state()
    ts = model.step(s, {"0": 1, "1": 1})                       # creates the staging of posg_model_step_host
    cfg = MCTSConfig(discount=0.95, search_time_limit=0.05, c=math.sqrt(2), truncated=False, action_selection="ucb",
                     state_belief_only=False, seed=0)
    for _ in range(2):
        planner = INTMCP.initialize(model, "0", cfg, nesting_level=1, search_policies=None, node_capacity=1 << 16,
                                    log_capacity=1 << 18, max_sims_per_level=64)
        planner.reset()
        planner.step(model.sample_initial_obs(s)["0"])
        planner.close()
        ts = model.step(ts.state, {"0": 0, "1": 2})
    assert ts.state is not None
