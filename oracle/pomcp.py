"""Scalar restatement of the reference's POMCP planner (oracle; test infrastructure
only) -- used as the CPU baseline of "POMCP simulations per decision" on the GPU box,
where /root/reference does not exist, and validated in the build container against the
unmodified reference planner (tests/test_oracle_planner.py).

Follows /root/reference/posggym_baselines/planning/mcts.py for the POMCP configuration
(pomcp.py:20-35: random other agents, state_belief_only=True; RandomSearchPolicy):
``_initial_update`` 175-227, ``_update`` 229-263, ``get_action`` 269-306, ``_simulate``
308-382, ``_rollout`` 405-452, ``ucb_action_selection`` 529-546,
``max_value_action_selection`` 583-600, ``_reinvigorate`` 651-700 with
``BeliefRejectionSampler._rejection_sample`` (belief.py:145-194), ``ActionNode.update``
(node.py:166-178) and ``MinMaxStats`` (utils.py:15-42).
"""
from __future__ import annotations

import math
import random
import time
from typing import Dict, List, Optional


class _ANode:
    __slots__ = ("action", "visits", "total_value", "value", "children")

    def __init__(self, action):
        self.action = action
        self.visits = 0
        self.total_value = 0.0
        self.value = 0.0
        self.children: Dict[object, "_ONode"] = {}

    def update(self, g: float):
        self.visits += 1
        self.total_value += g
        self.value += (g - self.value) / self.visits


class _ONode:
    __slots__ = ("obs", "t", "particles", "visits", "is_absorbing", "children", "parent")

    def __init__(self, parent, obs, t, visits=0):
        self.parent = parent
        self.obs = obs
        self.t = t
        self.particles: List[object] = []
        self.visits = visits
        self.is_absorbing = False
        self.children: Dict[int, _ANode] = {}


class OraclePOMCP:
    def __init__(self, model, agent_id: str, discount: float, search_time_limit: float, c: float,
                 epsilon: float = 0.01, extra_particles_prop: float = 1.0 / 16,
                 reinvigoration_sample_limit_factor: float = 4.0, step_limit: Optional[int] = None,
                 seed: Optional[int] = None):
        self.model, self.agent_id = model, agent_id
        self.discount, self.search_time_limit, self.c = discount, search_time_limit, c
        self.num_particles = math.ceil(100 * search_time_limit)
        self.extra_particles = math.ceil(self.num_particles * extra_particles_prop)
        self.depth_limit = 0 if discount == 0.0 else math.ceil(math.log(epsilon) / math.log(discount))
        self.sample_limit_factor = reinvigoration_sample_limit_factor
        self.step_limit = step_limit if step_limit is not None else (
            model.spec.max_episode_steps if model.spec is not None else float("inf"))
        self.actions = list(range(model.action_spaces[agent_id].n))
        self._rng = random.Random(seed)
        self._spaces = model.action_spaces
        self.step_statistics: Dict[str, float] = {}
        self.reset()

    def reset(self):
        self.root = _ONode(None, None, 0)
        self._last_action = None
        self._min, self._max = float("inf"), -float("inf")
        self.step_statistics = {"num_sims": 0, "search_time": 0.0, "search_depth": 0}

    # -- step ------------------------------------------------------------------------
    def step(self, obs):
        if self.root.is_absorbing:
            self.step_statistics = {k: float("nan") for k in self.step_statistics}
            return self._last_action
        self.update(self._last_action, obs)
        self._last_action = self.get_action()
        return self._last_action

    def update(self, action, obs):
        if self.root.is_absorbing:
            return
        if self.root.t == 0:
            node = _ONode(None, obs, 1)
            while len(node.particles) < self.num_particles + self.extra_particles:
                s = self.model.sample_initial_state()
                if self.model.sample_initial_obs(s)[self.agent_id] != obs:
                    continue
                node.particles.append(s)
            self.root = node
            return
        anode = self.root.children[action]
        child = anode.children.get(obs)
        if child is None:
            child = _ONode(anode, obs, self.root.t + 1)
            child.is_absorbing = self.root.is_absorbing
            anode.children[obs] = child
        if not child.is_absorbing:
            need = self.num_particles + self.extra_particles - len(child.particles)
            if need > 0:
                child.particles.extend(self._rejection_sample(action, obs, self.root.particles, need))
        child.parent = None
        self.root = child

    def _joint_action(self, ego_action):
        return {i: (ego_action if i == self.agent_id else self._spaces[i].sample()) for i in self.model.possible_agents}

    def _rejection_sample(self, action, obs, parent_particles, n):
        samples, rejected, attempts = [], [], 0
        limit = self.sample_limit_factor * n
        while len(samples) < n and attempts < limit:
            attempts += 1
            s = self._rng.choice(parent_particles)
            ts = self.model.step(s, self._joint_action(action))
            if ts.observations[self.agent_id] == obs:
                samples.append(ts.state)
            else:
                rejected.append(ts.state)
        if len(samples) < n:
            samples.extend(rejected[: n - len(samples)])
        return samples

    # -- search ------------------------------------------------------------------------
    def get_action(self, max_sims: Optional[int] = None):
        if self.root.is_absorbing:
            return self.actions[0]
        start = time.time()
        if not self.root.children:
            for a in self.actions:
                self.root.children[a] = _ANode(a)
        n_sims, max_depth = 0, 0
        while (time.time() - start < self.search_time_limit) if max_sims is None else (n_sims < max_sims):
            s = self._rng.choice(self.root.particles)
            _, d = self._simulate(s, self.root, 0)
            self.root.visits += 1
            max_depth = max(max_depth, d)
            n_sims += 1
        self.step_statistics = {"num_sims": n_sims, "search_time": time.time() - start, "search_depth": max_depth}
        best, best_v = [], -float("inf")
        for an in self.root.children.values():
            if an.value == best_v:
                best.append(an.action)
            elif an.value > best_v:
                best_v, best = an.value, [an.action]
        return random.choice(best)

    def _normalize(self, v):
        return (v - self._min) / (self._max - self._min) if self._max > self._min else v

    def _select(self, node: _ONode):
        if node.visits == 0:
            return random.choice(self.actions)
        log_n = math.log(node.visits)
        best, best_v = 0, -float("inf")
        for an in node.children.values():
            if an.visits == 0:
                return an.action
            v = self._normalize(an.value) + self.c * math.sqrt(log_n / an.visits)
            if v > best_v:
                best_v, best = v, an.action
        return best

    def _simulate(self, state, node: _ONode, depth: int):
        if depth > self.depth_limit or node.t > self.step_limit:
            return 0, depth
        if not node.children:
            for a in self.actions:
                node.children[a] = _ANode(a)
            return self._rollout(state, node.t, depth), depth
        a = self._select(node)
        ts = self.model.step(state, self._joint_action(a))
        ego_obs = ts.observations[self.agent_id]
        g = ts.rewards[self.agent_id]
        done = ts.terminations[self.agent_id] or ts.truncations[self.agent_id] or ts.all_done
        anode = node.children[a]
        child = anode.children.get(ego_obs)
        if child is None:
            child = _ONode(anode, ego_obs, node.t + 1, visits=1)
            anode.children[ego_obs] = child
        else:
            child.visits += 1
        child.is_absorbing = done
        child.particles.append(ts.state)
        max_depth = depth
        if not done:
            fut, max_depth = self._simulate(ts.state, child, depth + 1)
            g += self.discount * fut
        anode.update(g)
        self._max, self._min = max(self._max, anode.value), min(self._min, anode.value)
        return g, max_depth

    def _rollout(self, state, t, depth):
        ret, k = 0, 0
        while depth <= self.depth_limit and t <= self.step_limit:
            ts = self.model.step(state, {i: self._spaces[i].sample() for i in self.model.possible_agents})
            ret += self.discount**k * ts.rewards[self.agent_id]
            if ts.terminations[self.agent_id] or ts.truncations[self.agent_id] or ts.all_done:
                break
            state = ts.state
            depth += 1
            t += 1
            k += 1
        return ret
