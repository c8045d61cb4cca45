"""The C restatement of the Driving-v1 oracle (oracle/c/driving_oracle.c) against the
Python oracle's committed golden vectors.  CPU only."""
import numpy as np
import pytest

from tests.conftest import golden


def test_c_oracle_step_matches_golden(oracle_env):
    from oracle import native

    h = native.init_driving()
    g = golden("driving_step.npz")
    n = len(g["state"])
    state = np.ascontiguousarray(g["state"], dtype=np.int32)
    nxt = np.zeros_like(state)
    obs = np.zeros((n, 2, 22), np.int32)
    rew = np.zeros((n, 2), np.float64)
    flags = np.zeros(n, np.uint8)
    action = np.ascontiguousarray(g["action"])     # keep the buffers alive across the call
    order = np.ascontiguousarray(g["order"])
    h.drv_step(n, state.ctypes.data, action.ctypes.data, order.ctypes.data, nxt.ctypes.data, obs.ctypes.data,
               rew.ctypes.data, flags.ctypes.data)
    assert (nxt == g["next_state"]).all()
    assert (obs == g["obs"]).all()
    assert (rew.view(np.uint64) == g["reward"].view(np.uint64)).all()
    assert ((flags & 0x80).astype(bool) == g["all_done"]).all()
    assert ((flags & 1).astype(bool) == g["terminated"][:, 0]).all()


def test_c_oracle_rollouts_match_golden(oracle_env):
    from oracle import native

    h = native.init_driving()
    g = golden("driving_rollout.npz")
    n = len(g["particle"])
    ret = np.zeros(n, np.float64)
    steps = np.zeros(n, np.uint8)
    final = np.zeros((n, 2, 10), np.int32)
    h.drv_rollouts(int(g["seed"]), int(g["particle"][0]), n, 0, int(g["max_steps"]), float(g["discount"]),
                   ret.ctypes.data, steps.ctypes.data, final.ctypes.data)
    assert (ret.view(np.uint64) == g["ret"].view(np.uint64)).all()
    assert (steps == g["steps"]).all()
    assert (final == g["final_state"]).all()


def test_c_pp_oracle_matches_golden(oracle_env):
    from oracle import native

    h = native.init_pp()
    g = golden("pp_step.npz")
    n = len(g["state"])
    state, action, noise = (np.ascontiguousarray(g[k]) for k in ("state", "action", "noise"))
    nxt = np.zeros_like(state)
    obs = np.zeros((n, 2, 25), np.int8)
    rew = np.zeros((n, 2), np.float64)
    done = np.zeros(n, np.uint8)
    h.pp_step(n, state.ctypes.data, action.ctypes.data, noise.ctypes.data, nxt.ctypes.data, obs.ctypes.data,
              rew.ctypes.data, done.ctypes.data)
    assert (nxt == g["next_state"]).all()
    assert (obs == g["obs"]).all()
    assert (rew.view(np.uint64) == g["reward"].view(np.uint64)).all()
    assert (done.astype(bool) == g["all_done"]).all()
    g = golden("pp_rollout.npz")
    n = len(g["particle"])
    ret, steps, final = np.zeros(n, np.float64), np.zeros(n, np.uint8), np.zeros((n, 13), np.int8)
    h.pp_rollouts(int(g["seed"]), int(g["particle"][0]), n, int(g["ego"]), int(g["max_steps"]), float(g["discount"]),
                  ret.ctypes.data, steps.ctypes.data, final.ctypes.data)
    assert (ret.view(np.uint64) == g["ret"].view(np.uint64)).all()
    assert (steps == g["steps"]).all() and (final == g["final_state"]).all()


def test_c_lbf_oracle_matches_golden(oracle_env):
    from oracle import native

    h = native.init_lbf()
    g = golden("lbf_step.npz")
    n = len(g["state"])
    state, action = np.ascontiguousarray(g["state"]), np.ascontiguousarray(g["action"])
    nxt = np.zeros_like(state)
    obs = np.zeros((n, 2, 30), np.int8)
    rew = np.zeros((n, 2), np.float64)
    done = np.zeros(n, np.uint8)
    h.lbf_step(n, state.ctypes.data, action.ctypes.data, nxt.ctypes.data, obs.ctypes.data, rew.ctypes.data,
               done.ctypes.data)
    assert (nxt == g["next_state"]).all()
    assert (obs == g["obs"]).all()
    assert (rew.view(np.uint64) == g["reward"].view(np.uint64)).all()
    assert (done.astype(bool) == g["all_done"]).all()
    g = golden("lbf_rollout.npz")
    n = len(g["particle"])
    ret, steps, final = np.zeros(n, np.float64), np.zeros(n, np.uint8), np.zeros((n, 15), np.int8)
    h.lbf_rollouts(int(g["seed"]), int(g["particle"][0]), n, int(g["ego"]), int(g["max_steps"]),
                   float(g["discount"]), ret.ctypes.data, steps.ctypes.data, final.ctypes.data)
    assert (ret.view(np.uint64) == g["ret"].view(np.uint64)).all()
    assert (steps == g["steps"]).all() and (final == g["final_state"]).all()


def _py_hash(h, words):
    M = (1 << 64) - 1
    for w in words:
        h = ((h ^ int(w)) * 0x9E3779B97F4A7C15) & M
        h ^= h >> 32
    return h


@pytest.mark.parametrize("kind", ["driving", "pp", "lbf"])
def test_c_trace_matches_python_oracle(oracle_env, kind):
    """The C oracles' Philox-driven sweep (per-step packed words + rolling hash: what the GPU parity
    tests compare against at 2^16 .. 2^24 particles) replayed on the Python oracle, with the packed
    layouts of include/posg_b200.h written out independently by the host codecs."""
    from types import SimpleNamespace

    from oracle import native
    from oracle import rollout as R

    n, depth, seed, p0 = 48, 50, 20262, 1000
    parts = np.arange(n, dtype=np.uint64) + np.uint64(p0)
    if kind == "driving":
        from oracle.driving import DrivingModel
        from posggym_baselines_b200.envs.driving import BatchedDriving as B

        m, h = DrivingModel(), native.init_driving()
        codec = SimpleNamespace(num_agents=2, dest_coords=[tuple(c) for c in m.grid.dest_coords[0]])
        states = [m.initial_state_from_noise(p) for p in R.driving_initial_picks(m, seed, parts)]
    elif kind == "pp":
        from oracle.predator_prey import PredatorPreyModel
        from posggym_baselines_b200.envs.predator_prey import BatchedPredatorPrey as B

        m, h, codec = PredatorPreyModel(), native.init_pp(), None
        states = [m.initial_state_from_noise([int(v) for v in pk]) for pk in R.pp_initial_picks(m, seed, parts)]
    else:
        from oracle.lbf import LevelBasedForagingModel
        from posggym_baselines_b200.envs.lbf import BatchedLBF as B

        m, h = LevelBasedForagingModel(), native.init_lbf()
        codec = SimpleNamespace(food_coords=m.food_coords)
        states = [m.initial_state_from_noise([int(v) for v in pk]) for pk in R.lbf_initial_picks(m, seed, parts)]
    hashes, live, tr = native.trace(h, kind, seed, p0, n, depth, full=True)
    h2, live2, _ = native.trace(h, kind, seed, p0, n, depth, full=False, threads=3)
    assert (hashes == h2).all() and live == live2
    expect = [((int(p) + 1) * 0xD6E8FEB86659FD93) & ((1 << 64) - 1) for p in parts]
    expect_live = 0
    for k in range(depth):
        if kind == "driving":
            acts, order = R.driving_step_noise(seed, parts, k, 2)
        elif kind == "pp":
            acts, words = R.pp_step_noise(seed, parts, k)
        else:
            acts = R.lbf_step_noise(seed, parts, k)
        for i in range(n):
            ja = {"0": int(acts[i, 0]), "1": int(acts[i, 1])}
            if kind == "driving":
                expect_live += not all(v.dest_reached or v.crashed for v in states[i])
                ts = m.step_with_noise(states[i], ja, int(order[i]))
            elif kind == "pp":
                expect_live += not all(states[i].prey_caught)
                ts = m.step_with_noise(states[i], ja, [int(x) for x in words[i]])
            else:
                expect_live += any(f.level > 0 for f in states[i].food)
                ts = m.step_with_noise(states[i], ja, None)
            states[i] = ts.state
            sw = B.pack_state(codec, [ts.state]).view(np.uint32)[0]
            ow = np.concatenate([B.pack_obs(ts.observations[a]) for a in ("0", "1")])
            flags = (int(ts.terminations["0"]) | (int(ts.terminations["1"]) << 1) | (0x80 if ts.all_done else 0))
            rw = np.array([ts.rewards["0"], ts.rewards["1"]], np.float64)
            assert (tr["state"][k, i] == sw).all(), (k, i)
            assert (tr["obs"][k, i] == ow).all(), (k, i)
            assert (tr["rew"][k, i].view(np.uint64) == rw.view(np.uint64)).all(), (k, i)
            assert tr["flags"][k, i] == flags and (tr["act"][k, i] == acts[i]).all()
            expect[i] = _py_hash(expect[i], [*sw, *ow, *rw.view(np.uint64), flags | (int(acts[i, 0]) << 8) | (int(acts[i, 1]) << 16)])
    assert [int(x) for x in hashes] == expect
    assert live == expect_live
