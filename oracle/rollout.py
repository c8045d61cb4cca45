"""Scalar restatement of the reference's leaf rollout and of the Philox-driven
batch conventions (oracle; test infrastructure only).

``rollout`` follows ``MCTS._rollout`` (/root/reference/posggym_baselines/planning/
mcts.py:405-452) for ``state_belief_only=True`` with ``RandomSearchPolicy`` /
``RandomOtherAgentPolicy`` (search_policy.py:176-177, other_policy.py:150-151): the
loop bound ``depth <= depth_limit and t <= step_limit`` is pre-evaluated by the caller
into ``max_steps``; the uniform actions and the model's step noise come from the
Philox word (seed; particle, step0 + k; stream 0xA0; see driving_step_noise) instead of three host RNGs.
"""
from __future__ import annotations

import math
from typing import Sequence

import numpy as np

from oracle import philox as P


def driving_step_noise(seed: int, particles, step: int, num_agents: int = 2):
    """(actions [n, 2], order_idx [n]) exactly as the kernels derive them: ONE 32-bit
    Philox word per env-step -- block = counter (particle >> 2, step, stream 0xA0), word
    = particle & 3; idx = floor(word * 50 / 2**32) enumerates the 50 (order, a0, a1)
    combinations: order = idx % 2, a0 = (idx // 2) % 5, a1 = idx // 10."""
    assert num_agents == 2
    particles = np.asarray(particles, dtype=np.uint64)
    w = P.philox_block(seed, particles >> np.uint64(2), step, P.STREAM_STEP)
    which = (particles & np.uint64(3)).astype(np.int64)
    word = np.choose(which, w)
    idx = P.uniform_below(word, 50)
    order = idx & 1
    r = idx >> 1
    a1 = r // 5
    a0 = r - 5 * a1
    return np.stack([a0, a1], axis=-1), order


def driving_initial_picks(model, seed: int, particles: Sequence[int]):
    """Start/destination picks per particle as the kernels derive them from Philox."""
    na = len(model.possible_agents)
    assert na == 2
    particles = np.asarray(particles, dtype=np.uint64)
    w = P.philox_block(seed, particles, 0, P.STREAM_INIT)
    words = [(w[0], w[1]), (w[2], w[3])]  # agent 0: words x, y; agent 1: words z, w
    out = []
    for p in range(len(particles)):
        picks = []
        for i in range(na):
            for word in words[i]:
                bound = model.initial_noise_bounds(picks)
                picks.append(int(P.uniform_below(word[p], bound)))
        out.append(picks)
    return out


def rollout(model, state, ego: str, max_steps: int, discount: float, seed: int, particle: int,
            step0: int = 0):
    """Returns (discounted ego return, model steps executed, final state)."""
    agent_return = 0
    k = 0
    na = len(model.possible_agents)
    while k < max_steps:
        acts, order = driving_step_noise(seed, np.asarray([particle], dtype=np.uint64), step0 + k, na)
        joint_action = {i: int(acts[0, j]) for j, i in enumerate(model.possible_agents)}
        ts = model.step_with_noise(state, joint_action, int(order[0]))
        agent_return += discount**k * ts.rewards[ego]
        state = ts.state
        k += 1
        if ts.terminations[ego] or ts.truncations[ego] or ts.all_done:
            break
    return float(agent_return), k, state


# ---- PredatorPrey-v0 conventions (PPSim in posggym_baselines_b200/csrc/pp.cuh) ----------
def pp_step_noise(seed: int, particles, step: int):
    """(actions [n, 2], prey words [n, 3]): HALF a Philox block per (particle, step).  Particles g and g + 32 (bit 5
    of g clear) share the block with counter ((g >> 6) << 5 | (g & 31), step): the first takes the words (a, b) =
    (x, y), the second (z, w).  Prey 0 uses a, prey 1 a * 0x10001 mod 2**32, prey 2 b; the joint action comes from
    b * 0x10001 mod 2**32: idx = floor(word * 25 / 2**32), a0 = idx % 5, a1 = idx // 5."""
    particles = np.asarray(particles, dtype=np.uint64)
    counter = ((particles >> np.uint64(6)) << np.uint64(5)) | (particles & np.uint64(31))
    w = P.philox_block(seed, counter, step, P.STREAM_STEP)
    hi = ((particles >> np.uint64(5)) & np.uint64(1)).astype(bool)
    a = np.where(hi, w[2], w[0]).astype(np.uint64)
    b = np.where(hi, w[3], w[1]).astype(np.uint64)
    m32 = np.uint64(0xFFFFFFFF)
    a2 = (a * np.uint64(0x10001)) & m32
    b2 = (b * np.uint64(0x10001)) & m32
    idx = P.uniform_below(b2.astype(np.uint32), 25)
    a1 = idx // 5
    a0 = idx - 5 * a1
    return np.stack([a0, a1], axis=-1), np.stack([a, a2, b], axis=-1).astype(np.uint32)


def pp_initial_picks(model, seed: int, particles):
    """Picks as the kernel derives them: block (particle, 0, 0xB0) words x, y -> predators;
    block (particle, 0, 0xB1) words x, y, z -> prey."""
    particles = np.asarray(particles, dtype=np.uint64)
    a = P.philox_block(seed, particles, 0, P.STREAM_INIT)
    b = P.philox_block(seed, particles, 0, P.STREAM_INIT2)
    bounds = model.initial_noise_bounds()
    words = [a[0], a[1], b[0], b[1], b[2]]
    return np.stack([P.uniform_below(w, bd) for w, bd in zip(words, bounds)], axis=-1)


def pp_rollout(model, state, ego: str, max_steps: int, discount: float, seed: int, particle: int, step0: int = 0):
    agent_return, k = 0, 0
    while k < max_steps:
        acts, words = pp_step_noise(seed, np.asarray([particle], dtype=np.uint64), step0 + k)
        ts = model.step_with_noise(state, {"0": int(acts[0, 0]), "1": int(acts[0, 1])}, [int(x) for x in words[0]])
        agent_return += discount**k * ts.rewards[ego]
        state = ts.state
        k += 1
        if ts.terminations[ego] or ts.truncations[ego] or ts.all_done:
            break
    return float(agent_return), k, state


# ---- LevelBasedForaging-v3 conventions (LBFSim in posggym_baselines_b200/csrc/lbf.cuh) ----
def lbf_step_noise(seed: int, particles, step: int):
    """actions [n, 2]: ONE 32-bit Philox word per env-step, as for Driving -- block = counter
    (particle >> 2, step, stream 0xA0), word = particle & 3; idx = floor(word * 36 / 2**32),
    a0 = idx % 6, a1 = idx // 6.  The LBF transition itself takes no noise."""
    particles = np.asarray(particles, dtype=np.uint64)
    w = P.philox_block(seed, particles >> np.uint64(2), step, P.STREAM_STEP)
    word = np.choose((particles & np.uint64(3)).astype(np.int64), w)
    idx = P.uniform_below(word, 36)
    a1 = idx // 6
    a0 = idx - 6 * a1
    return np.stack([a0, a1], axis=-1)


def lbf_initial_picks(model, seed: int, particles):
    """Picks as the kernel derives them: block (particle, 0, 0xB0) words x, y -> player
    levels, z, w -> foods 0, 1; block (.., 0xB1) -> foods 2..5; block (.., 0xB2) words x, y ->
    foods 6, 7."""
    particles = np.asarray(particles, dtype=np.uint64)
    a = P.philox_block(seed, particles, 0, P.STREAM_INIT)
    b = P.philox_block(seed, particles, 0, P.STREAM_INIT2)
    c = P.philox_block(seed, particles, 0, 0xB2)
    p0 = P.uniform_below(a[0], model.max_agent_level)
    p1 = P.uniform_below(a[1], model.max_agent_level)
    bound = np.ones_like(p0) if model.force_coop else p0 + p1 + 2
    foods = [P.uniform_below(w, bound) for w in (a[2], a[3], b[0], b[1], b[2], b[3], c[0], c[1])]
    return np.stack([p0, p1, *foods], axis=-1)


def lbf_rollout(model, state, ego: str, max_steps: int, discount: float, seed: int, particle: int, step0: int = 0):
    agent_return, k = 0, 0
    while k < max_steps:
        acts = lbf_step_noise(seed, np.asarray([particle], dtype=np.uint64), step0 + k)
        ts = model.step_with_noise(state, {"0": int(acts[0, 0]), "1": int(acts[0, 1])}, None)
        agent_return += discount**k * ts.rewards[ego]
        state = ts.state
        k += 1
        if ts.terminations[ego] or ts.truncations[ego] or ts.all_done:
            break
    return float(agent_return), k, state
