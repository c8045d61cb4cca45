"""The C restatement of the Driving-v1 oracle (oracle/c/driving_oracle.c) against the
Python oracle's committed golden vectors.  CPU only."""
import numpy as np

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
