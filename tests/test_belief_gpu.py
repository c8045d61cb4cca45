"""Particle-belief kernels against a sequential replay on the CPU oracle: the reference's
rejection loops are order-sensitive ("first N accepted in attempt order, then the first
missing rejects in attempt order", planning/belief.py:145-194; "until N match the first
observation", planning/mcts.py:186-194).  With the Philox draws replayed, the CUDA output
must equal the sequential loop's output element for element."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    from posggym_baselines_b200.envs.driving import BatchedDriving

    return BatchedDriving()


def obs_key(env, o):
    return np.ascontiguousarray(env.pack_obs(o), dtype=np.uint64)


@pytest.mark.parametrize("need,ego,action", [(11, 0, 1), (200, 1, 3), (2125, 0, 0)])
def test_reinvigoration_matches_sequential_rejection_loop(env, oracle_env, need, ego, action):
    from oracle import philox as P
    from oracle.driving import DrivingModel
    from posggym_baselines_b200 import _lib

    m = DrivingModel()
    seed, draw_id, n_parent = 2024, 777, 64
    # parent belief: states a few random steps into an episode
    parent = env.sample_initial_state(n_parent, seed=5)
    for k in range(6):
        parent = env.step(parent, seed=5, step_idx=k).state
    parent_cpu = env.unpack_state(parent)
    to_oracle = lambda s: tuple(type(m.initial_state_from_noise([0, 0, 0, 0])[0])(*v) for v in s)  # noqa: E731
    max_attempts = 4 * need

    def draw(att):
        w = P.philox_block(seed, np.array([draw_id + att], dtype=np.uint64), 0xFFFE, P.STREAM_BELIEF)
        pi = (int(w[0][0]) * n_parent) >> 32
        idx = (int(w[1][0]) * 50) >> 32
        order, r = idx & 1, idx >> 1
        a1 = r // 5
        a0 = r - 5 * a1
        acts = [a0, a1]
        acts[ego] = action
        return pi, acts, order

    # the target observation: what attempt 0 produces (so that some attempts are accepted)
    pi, acts, order = draw(0)
    ts = m.step_with_noise(to_oracle(parent_cpu[pi]), {"0": acts[0], "1": acts[1]}, order)
    target = ts.observations[str(ego)]
    accepted, rejected = [], []
    for att in range(max_attempts):
        if len(accepted) >= need:
            break
        pi, acts, order = draw(att)
        ts = m.step_with_noise(to_oracle(parent_cpu[pi]), {"0": acts[0], "1": acts[1]}, order)
        (accepted if ts.observations[str(ego)] == target else rejected).append(ts.state)
    expect = accepted[:need] + rejected[: max(0, need - len(accepted))]

    out = torch.empty((need, 2), dtype=torch.int32, device=env.device)
    rej = torch.empty((need, 2), dtype=torch.int32, device=env.device)
    res = np.zeros(3, np.int64)
    key = obs_key(env, target)
    _lib.check(_lib.lib().posg_belief_reinvigorate(
        env._handle, _lib.ptr(parent), n_parent, ego, action, _lib.ptr(key), need, max_attempts, seed, draw_id,
        _lib.ptr(out), _lib.ptr(rej), _lib.ptr(res), None))
    got = env.unpack_state(out[: int(res[2])])
    assert int(res[0]) == min(need, len(accepted)) and int(res[2]) == len(expect)
    assert [tuple(map(tuple, s)) for s in got] == [tuple(map(tuple, s)) for s in expect]
    assert 0 < len(accepted)   # the case exercises both branches when need is large
    if need >= 200:
        assert len(rejected) > 0


def test_initial_belief_matches_sequential_rejection_loop(env, oracle_env):
    from oracle import rollout as R
    from oracle.driving import DrivingModel
    from posggym_baselines_b200 import _lib

    m = DrivingModel()
    seed, draw_id, need, ego = 99, 12345, 300, 1
    picks = R.driving_initial_picks(m, seed, np.arange(draw_id, draw_id + 40_000, dtype=np.uint64))
    target = m.sample_initial_obs(m.initial_state_from_noise(picks[3]))[str(ego)]
    expect = []
    for pk in picks:
        s = m.initial_state_from_noise(pk)
        if m.sample_initial_obs(s)[str(ego)] == target:
            expect.append(s)
            if len(expect) == need:
                break
    assert len(expect) == need
    out = torch.empty((need, 2), dtype=torch.int32, device=env.device)
    res = np.zeros(3, np.int64)
    key = obs_key(env, target)
    _lib.check(_lib.lib().posg_belief_initial(env._handle, ego, _lib.ptr(key), need, 1 << 20, seed, draw_id,
                                              _lib.ptr(out), _lib.ptr(res), None))
    assert int(res[0]) == need
    got = env.unpack_state(out)
    assert [tuple(map(tuple, s)) for s in got] == [tuple(map(tuple, s)) for s in expect]
