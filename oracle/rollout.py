"""Scalar restatement of the reference's leaf rollout and of the Philox-driven
batch conventions (oracle; test infrastructure only).

``rollout`` follows ``MCTS._rollout`` (/root/reference/posggym_baselines/planning/
mcts.py:405-452) for ``state_belief_only=True`` with ``RandomSearchPolicy`` /
``RandomOtherAgentPolicy`` (search_policy.py:176-177, other_policy.py:150-151): the
loop bound ``depth <= depth_limit and t <= step_limit`` is pre-evaluated by the caller
into ``max_steps``; the uniform actions and the model's step noise come from the
Philox block (seed; particle, step0 + k; stream 0xA0) instead of three host RNGs.
"""
from __future__ import annotations

import math
from typing import Sequence

import numpy as np

from oracle import philox as P


def driving_step_noise(seed: int, particles, step: int, num_agents: int):
    """(actions [n, A], order_idx [n]) exactly as the kernels derive them."""
    w = P.philox_block(seed, particles, step, P.STREAM_STEP)
    acts = [P.uniform_below(w[0], 5)]
    if num_agents > 1:
        acts.append(P.uniform_below(w[1], 5))
    order = P.uniform_below(w[2], math.factorial(num_agents))
    if num_agents > 2:
        w2 = P.philox_block(seed, particles, step, P.STREAM_STEP2)
        acts.append(P.uniform_below(w2[0], 5))
        if num_agents > 3:
            acts.append(P.uniform_below(w2[1], 5))
    return np.stack(acts, axis=-1), order


def driving_initial_picks(model, seed: int, particles: Sequence[int]):
    """Start/destination picks per particle as the kernels derive them from Philox."""
    na = len(model.possible_agents)
    particles = np.asarray(particles, dtype=np.uint64)
    w = P.philox_block(seed, particles, 0, P.STREAM_INIT)
    w2 = P.philox_block(seed, particles, 0, P.STREAM_INIT2) if na > 2 else w
    words = []
    for i in range(na):
        src = w if i < 2 else w2
        words.append((src[0 if i % 2 == 0 else 2], src[1 if i % 2 == 0 else 3]))
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
