"""SURVEY 8(d) config 2's parity gate for all three models: the random-action ``model.step`` sweep
(Philox initial states, Philox joint actions, every particle stepped ``depth`` times) through the
C-ABI against the compiled CPU oracles (oracle/c/*_oracle.c, pinned to the Python oracles' goldens in
tests/test_oracle_native.py):

* FULL per-step trace -- packed state, both observations, both float64 rewards, flags, actions of
  every step -- compared bit for bit at 2^16 particles x 50 steps (3.3 M (state, action, noise)
  triples per model);
* the 64-bit rolling hash of the same words (posg_trace_hash) per particle at 2^24 particles x 50
  steps for Driving-v1 (8.4e8 triples) and 2^21 x 50 for PredatorPrey-v0 / LevelBasedForaging-v3
  (1.0e8 triples each), plus the executed-env-step count.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

DEPTH = 50


def make(kind):
    from oracle import native

    if kind == "driving":
        from posggym_baselines_b200.envs.driving import BatchedDriving

        return BatchedDriving(), native.init_driving()
    if kind == "pp":
        from posggym_baselines_b200.envs.predator_prey import BatchedPredatorPrey

        return BatchedPredatorPrey(), native.init_pp()
    from posggym_baselines_b200.envs.lbf import BatchedLBF

    return BatchedLBF(), native.init_lbf()


def gpu_sweep(env, n, seed, off, per_step=None):
    from posggym_baselines_b200.envs.base import trace_hash

    state = env.sample_initial_state(n, seed=seed, particle_offset=off)
    hashes = torch.empty(n, dtype=torch.int64, device=env.device)
    trace_hash(hashes, None, init=True, particle_offset=off)
    live = torch.zeros(1, dtype=torch.int64, device=env.device)
    cur, out = state, None
    for k in range(DEPTH):
        out = env.step(cur, seed=seed, step_idx=k, particle_offset=off, live_count=live,
                       out=out if per_step is None else None)
        trace_hash(hashes, out)
        if per_step is not None:
            per_step(k, out)
            cur, out = out.state, None
        else:
            # ping-pong the state buffers, reuse everything else
            cur, out = out.state, out._replace(state=cur if k else torch.empty_like(cur))
    torch.cuda.synchronize()
    return hashes.cpu().numpy().view(np.uint64), int(live.item())


@pytest.mark.parametrize("kind", ["driving", "pp", "lbf"])
def test_full_trace_bit_exact_at_65536_particles(oracle_env, kind):
    from oracle import native

    env, h = make(kind)
    n, seed, off = 1 << 16, 8086, 3 << 20
    c_hash, c_live, tr = native.trace(h, kind, seed, off, n, DEPTH, full=True)
    mismatches = []

    def check(k, out):
        st = out.state.cpu().numpy().view(np.uint32).reshape(n, -1)
        ob = out.obs.cpu().numpy().view(np.uint64).reshape(n, -1)
        rw = out.rewards.cpu().numpy().view(np.uint64)
        fl, ac = out.flags.cpu().numpy(), out.actions.cpu().numpy()
        for name, got, want in (("state", st, tr["state"][k]), ("obs", ob, tr["obs"][k]),
                                ("rew", rw, tr["rew"][k].view(np.uint64)), ("flags", fl, tr["flags"][k]),
                                ("act", ac, tr["act"][k])):
            if not np.array_equal(got, want):
                bad = np.nonzero((got != want).reshape(n, -1).any(1))[0]
                mismatches.append((k, name, len(bad), bad[:5].tolist()))

    g_hash, g_live = gpu_sweep(env, n, seed, off, per_step=check)
    assert not mismatches, mismatches[:10]
    assert np.array_equal(g_hash, c_hash)
    assert g_live == c_live


@pytest.mark.parametrize("kind,log2n", [("driving", 24), ("pp", 21), ("lbf", 21)])
def test_rolling_hash_at_scale(oracle_env, kind, log2n):
    from oracle import native

    env, h = make(kind)
    n, seed, off = 1 << log2n, 20261019, 1 << 30
    c_hash, c_live, _ = native.trace(h, kind, seed, off, n, DEPTH)
    g_hash, g_live = gpu_sweep(env, n, seed, off)
    bad = np.nonzero(g_hash != c_hash)[0]
    assert len(bad) == 0, (len(bad), bad[:10].tolist())
    assert g_live == c_live
    # a checksum of checksums for the record (printed with -s)
    print(kind, n, "xor of hashes", hex(int(np.bitwise_xor.reduce(g_hash))), "live env-steps", g_live)
