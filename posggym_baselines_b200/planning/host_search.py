"""Plugin path of the planner: arbitrary ``SearchPolicy`` / ``OtherAgentPolicy`` objects.

The device search (planning/mcts.py + csrc/search.cuh) executes ``RandomSearchPolicy`` /
``RandomOtherAgentPolicy`` inside its kernels.  Any other plugin is a Python object with state that only
its scalar methods (``sample_action`` / ``get_pi`` / ``get_value`` / ``get_next_state``,
/root/reference/posggym_baselines/planning/search_policy.py:16-155, other_policy.py:10-129) can advance,
so the tree descent that calls them runs on the host -- the algorithm of the reference engine
(planning/mcts.py:159-263 update, 269-306 get_action, 308-382 _simulate, 384-452 _evaluate / _rollout,
494-600 action selection, 651-700 reinvigoration; belief.py:145-194) written against the same model
surface -- while everything that is model arithmetic stays on the device and, where the algorithm allows
it, batched:

* the initial belief (``MCTS._initial_update``'s rejection loop) is ONE ``posg_belief_initial`` launch;
* reinvigoration steps a whole round of rejection-sampling attempts with ONE ``model.step`` launch
  (parents and joint actions chosen on the host, in the order the sequential loop would, so the accepted /
  padded particles are the ones it would produce for the same draws);
* single model steps inside a simulation go through the facade's scalar ``model.step``;
* both kinds of step are ``posg_model_step_host`` calls: host rows staged through one pinned block, one copy in,
  the step kernel, one copy out, one synchronisation -- no tensor objects on the way.

Particles carry ``(state, joint history, other agents' policy states, t)`` unless
``config.state_belief_only``.
"""
from __future__ import annotations

import math
import random
import time
from typing import Any, Dict, List, NamedTuple, Optional

import numpy as np
import torch


class HPS(NamedTuple):
    """``HistoryPolicyState`` (planning/belief.py:12-44)."""

    state: Any
    history: Optional[tuple]      # ((joint action, joint obs), ...) or None
    policy_state: Optional[Dict[str, Any]]
    t: int


class _Belief:
    def __init__(self, rng: random.Random):
        self.particles: List[HPS] = []
        self._rng = rng

    def size(self) -> int:
        return len(self.particles)

    def add_particle(self, p: HPS):
        self.particles.append(p)

    def sample(self) -> HPS:
        return self._rng.choice(self.particles)


class _ActionNode:
    __slots__ = ("parent", "action", "prob", "visits", "total_value", "value", "agg", "children")

    def __init__(self, parent, action, prob):
        self.parent, self.action, self.prob = parent, action, prob
        self.visits, self.total_value, self.value, self.agg = 0, 0.0, 0.0, 0.0
        self.children: Dict[Any, "_ObsNode"] = {}

    def update(self, ret: float):   # node.py:166-178 (running mean + Welford aggregate)
        self.visits += 1
        self.total_value += ret
        delta = ret - self.value
        self.value += delta / self.visits
        self.agg += delta * (ret - self.value)

    @property
    def variance(self) -> float:
        return self.agg / self.visits if self.visits else 0.0


class _ObsNode:
    __slots__ = ("parent", "obs", "t", "belief", "action_probs", "search_policy_state", "visits", "is_absorbing",
                 "children")

    def __init__(self, parent, obs, t, belief, action_probs, search_policy_state, visits=0):
        self.parent, self.obs, self.t, self.belief = parent, obs, t, belief
        self.action_probs, self.search_policy_state = action_probs, search_policy_state or {}
        self.visits, self.is_absorbing = visits, False
        self.children: Dict[Any, _ActionNode] = {}

    def add_child(self, action) -> _ActionNode:
        node = _ActionNode(self, action, self.action_probs.get(action, 0.0))
        self.children[action] = node
        return node


class _MinMax:
    def __init__(self, known):
        self.minimum, self.maximum = (known.min, known.max) if known else (float("inf"), -float("inf"))

    def update(self, v: float):
        self.minimum, self.maximum = min(self.minimum, v), max(self.maximum, v)

    def normalize(self, v: float) -> float:
        return (v - self.minimum) / (self.maximum - self.minimum) if self.maximum > self.minimum else v


class HostSearch:
    """The engine behind ``MCTS`` when a plugin is not one of the in-kernel Random policies."""

    def __init__(self, planner):
        self.p = planner
        self.model, self.agent_id, self.config = planner.model, planner.agent_id, planner.config
        self.other_agent_policies, self.search_policy = planner.other_agent_policies, planner.search_policy
        self.action_space = planner.action_space
        self.step_limit = planner.step_limit
        self._rng = random.Random(self.config.seed)
        self._select = {"pucb": self._pucb, "ucb": self._ucb}.get(self.config.action_selection, self._min_visit)
        self.reset()

    # ------------------------------------------------------------------ plumbing
    @property
    def stats(self):
        return self.p.step_statistics

    def reset(self):
        self.min_max = _MinMax(self.config.known_bounds)
        self.root = _ObsNode(None, None, 0, _Belief(self._rng), {None: 1.0}, self.search_policy.get_initial_state())

    def _policy_next(self, action, obs, policy, state):
        t0 = time.time()
        nxt = policy.get_next_state(action, obs, state)
        self.stats["inference_time"] += time.time() - t0
        self.stats["policy_calls"] += 1
        return nxt

    def _others_next(self, joint_action, joint_obs, pi_state):
        return {i: self._policy_next(joint_action[i], joint_obs[i], self.other_agent_policies[i], pi_state[i])
                for i in self.model.possible_agents if i != self.agent_id and i in joint_action}

    def _joint_action(self, hps: HPS, ego_action):
        out = {}
        for i in self.model.possible_agents:
            if i == self.agent_id:
                out[i] = ego_action
            elif self.config.state_belief_only:
                out[i] = self.other_agent_policies[i].sample_action({})
            else:
                out[i] = self.other_agent_policies[i].sample_action(hps.policy_state[i])
        return out

    def _add_obs_node(self, parent: _ActionNode, obs, init_visits=0, sim_policy=None) -> _ObsNode:
        """``sim_policy`` is the search policy of the simulation that creates the node; None on the update path."""
        sp_state = self._policy_next(parent.action, obs, self.search_policy, parent.parent.search_policy_state)
        node = _ObsNode(parent, obs, parent.parent.t + 1, _Belief(self._rng), self._node_prior(sp_state, sim_policy),
                        sp_state, init_visits)
        parent.children[obs] = node
        return node

    # -- the places where a typed planner (planning/potmmcp.py) differs from MCTS
    def _sim_policy(self, hps: HPS):
        """Search policy of one simulation (mcts.py:285-290: the planner's own)."""
        return self.search_policy

    def _node_prior(self, sp_state, sim_policy):
        return self.search_policy.get_pi(sp_state)

    def _leaf_state(self, node: _ObsNode, sim_policy):
        return node.search_policy_state

    def _revisit(self, node: _ObsNode, child: _ObsNode, sim_policy):
        """An existing child is reached again: moving average of the prior (mcts.py:359-366)."""
        child.visits += 1
        pi = sim_policy.get_pi(child.search_policy_state)
        for a, pr in pi.items():
            child.action_probs[a] += (pr - child.action_probs[a]) / child.visits

    def _advance(self, hps: HPS, joint_action, ts) -> HPS:
        if self.config.state_belief_only:
            return HPS(ts.state, None, None, hps.t + 1)
        hist = hps.history + ((dict(joint_action), dict(ts.observations)),)
        return HPS(ts.state, hist, self._others_next(joint_action, ts.observations, hps.policy_state), hps.t + 1)

    # ------------------------------------------------------------------ update
    def update(self, action, obs):
        if self.root.is_absorbing:
            return
        if self.root.t == 0:
            self._initial_update(obs)
        else:
            self._update(action, obs)

    def _initial_update(self, init_obs):
        cfg, model, b = self.config, self.model, self.model.batched
        anode = self.root.add_child(None)
        node = self._add_obs_node(anode, init_obs, 0)
        need = cfg.num_particles + cfg.extra_particles
        states = self.p._device_initial_belief(init_obs, need)        # ONE launch: the rejection loop
        packed = states
        joint_obs_all = b.unpack_obs(b.observe(packed))
        unpacked = b.unpack_state(packed)
        ids = list(model.possible_agents)
        for k in range(need):
            joint_obs = {i: joint_obs_all[k][n] for n, i in enumerate(ids)}
            if cfg.state_belief_only:
                hist, pstate = None, None
            else:
                hist = ((dict.fromkeys(ids), dict(joint_obs)),)
                pstate = {j: self.other_agent_policies[j].sample_initial_state() for j in ids if j != self.agent_id}
                pstate = self._others_next({i: None for i in ids}, joint_obs, pstate)
            node.belief.add_particle(HPS(unpacked[k], hist, pstate, 1))
        self.root = node
        node.parent = None

    def _update(self, action, obs):
        anode = self.root.children.get(action)
        if anode is None:
            if not self.root.is_absorbing:
                raise AssertionError(f"root has no child node for action={action}")
            anode = self.root.add_child(action)
        node = anode.children.get(obs)
        if node is None:
            node = self._add_obs_node(anode, obs, 0)
            node.is_absorbing = self.root.is_absorbing
        if not node.is_absorbing:
            self._reinvigorate(node, action, obs)
        self.root = node
        node.parent = None

    def _reinvigorate(self, node: _ObsNode, action, obs):
        """BeliefRejectionSampler._rejection_sample with use_rejected_samples=True, a round of attempts per
        ``model.step`` launch."""
        t0 = time.time()
        cfg, b = self.config, self.model.batched
        need = cfg.num_particles + cfg.extra_particles - node.belief.size()
        if need > 0:
            parent = node.parent.parent.belief
            limit = int(cfg.reinvigoration_sample_limit_factor * need)
            ids = list(self.model.possible_agents)
            ego = ids.index(self.agent_id)
            want = b.pack_obs(obs)
            accepted, rejected, attempts = [], [], 0
            while len(accepted) < need and attempts < limit:
                n = min(limit - attempts, max(32, 2 * (need - len(accepted))))
                parents = [parent.sample() for _ in range(n)]
                actions = [self._joint_action(h, action) for h in parents]
                act = np.array([[int(a[i]) for i in ids] for a in actions], dtype=np.uint8)
                nst, nob, rew, flags = b.step_host(b.pack_state([h.state for h in parents]), act,
                                                   self.model.host_step_noise(n))
                keys = np.ascontiguousarray(nob.reshape(n, len(ids), -1)[:, ego, :]).view(np.uint64)
                match = (keys == want.reshape(1, -1)).all(axis=1)
                nstates, nobs = b.unpack_state(nst), b.unpack_obs(nob)
                for k in range(n):
                    if len(accepted) >= need:      # the sequential loop stops here: later attempts never happened
                        break
                    attempts += 1
                    ts = _Step(nstates[k], {i: nobs[k][m] for m, i in enumerate(ids)}, rew[k], flags[k], ids)
                    nxt = self._advance(parents[k], actions[k], ts)
                    (accepted if match[k] else rejected).append(nxt)
            if len(accepted) < need:
                accepted.extend(rejected[: need - len(accepted)])
            for hps in accepted:
                node.belief.add_particle(hps)
        self.stats["reinvigoration_time"] += time.time() - t0

    # ------------------------------------------------------------------ search
    def get_action(self):
        if self.root.is_absorbing:
            return self.action_space[0]
        cfg = self.config
        start = time.time()
        if not self.root.children:
            for a in self.action_space:
                self.root.add_child(a)
        n_sims, max_depth = 0, 0
        max_sims = self.p.max_sims
        while (time.time() - start < cfg.search_time_limit) if max_sims is None else (n_sims < max_sims):
            hps = self.root.belief.sample()
            _, depth = self._simulate(hps, self.root, 0, self._sim_policy(hps))
            self.root.visits += 1
            max_depth = max(max_depth, depth)
            n_sims += 1
        self.stats.update({"search_time": time.time() - start, "search_depth": max_depth, "num_sims": n_sims,
                           "min_value": self.min_max.minimum, "max_value": self.min_max.maximum})
        if cfg.action_selection == "pucb":
            return self._max_visit(self.root)
        return self._max_value(self.root)

    def _simulate(self, hps: HPS, node: _ObsNode, depth: int, sim_policy):
        cfg = self.config
        if depth > cfg.depth_limit or node.t > self.step_limit:
            return 0.0, depth
        if not node.children:
            for a in self.action_space:
                node.add_child(a)
            return self._evaluate(hps, depth, sim_policy, self._leaf_state(node, sim_policy)), depth
        ego_action = self._select(node)
        joint_action = self._joint_action(hps, ego_action)
        ts = self.model.step(hps.state, joint_action)
        ego_obs = ts.observations[self.agent_id]
        ret = ts.rewards[self.agent_id]
        done = ts.terminations[self.agent_id] or ts.truncations[self.agent_id] or ts.all_done
        nxt = self._advance(hps, joint_action, ts)
        anode = node.children[ego_action]
        child = anode.children.get(ego_obs)
        if child is not None:
            self._revisit(node, child, sim_policy)
        else:
            child = self._add_obs_node(anode, ego_obs, 1, sim_policy)
        child.is_absorbing = done
        child.belief.add_particle(nxt)
        max_depth = depth
        if not done:
            future, max_depth = self._simulate(nxt, child, depth + 1, sim_policy)
            ret += cfg.discount * future
        anode.update(ret)
        self.min_max.update(anode.value)
        return ret, max_depth

    def _evaluate(self, hps: HPS, depth: int, policy, sp_state) -> float:
        t0 = time.time()
        cfg = self.config
        if cfg.truncated:
            try:
                v = policy.get_value(sp_state)
            except NotImplementedError:
                if not cfg.use_rollout_if_no_value:
                    raise
                v = self._rollout(hps, depth, policy, sp_state)
        else:
            v = self._rollout(hps, depth, policy, sp_state)
        self.stats["evaluation_time"] += time.time() - t0
        return v

    def _rollout(self, hps: HPS, depth: int, policy, sp_state) -> float:
        cfg, ret, k = self.config, 0, 0
        while depth <= cfg.depth_limit and hps.t <= self.step_limit:
            ego_action = policy.sample_action(sp_state)
            joint_action = self._joint_action(hps, ego_action)
            ts = self.model.step(hps.state, joint_action)
            ret += cfg.discount**k * ts.rewards[self.agent_id]
            if ts.terminations[self.agent_id] or ts.truncations[self.agent_id] or ts.all_done:
                break
            hps = self._advance(hps, joint_action, ts)
            sp_state = self._policy_next(ego_action, ts.observations[self.agent_id], policy, sp_state)
            depth += 1
            k += 1
        return ret

    # ------------------------------------------------------------------ action selection
    def _pucb(self, node: _ObsNode):
        if node.visits == 0:
            return random.choices(list(node.action_probs), weights=list(node.action_probs.values()), k=1)[0]
        frac, noise = self.config.pucb_exploration_fraction, 1.0 / len(self.action_space)
        sqrt_n = math.sqrt(node.visits)
        best, best_a = -float("inf"), 0
        for an in node.children.values():
            prior = node.action_probs[an.action] * (1 - frac) + frac * noise
            v = (self.min_max.normalize(an.value) if an.visits > 0 else 0) + self.config.c * prior * (sqrt_n / (1 + an.visits))
            if v > best:
                best, best_a = v, an.action
        return best_a

    def _ucb(self, node: _ObsNode):
        if node.visits == 0:
            return random.choice(self.action_space)
        log_n = math.log(node.visits)
        best, best_a = -float("inf"), 0
        for an in node.children.values():
            if an.visits == 0:
                return an.action
            v = self.min_max.normalize(an.value) + self.config.c * math.sqrt(log_n / an.visits)
            if v > best:
                best, best_a = v, an.action
        return best_a

    def _min_visit(self, node: _ObsNode):
        if node.visits == 0:
            return random.choice(self.action_space)
        return min(node.children.values(), key=lambda an: an.visits).action

    def _max_visit(self, node: _ObsNode):
        if node.visits == 0:
            return random.choice(self.action_space)
        top = max(an.visits for an in node.children.values())
        return random.choice([an.action for an in node.children.values() if an.visits == top])

    def _max_value(self, node: _ObsNode):
        if not node.children:
            return random.choice(self.action_space)
        top = max(an.value for an in node.children.values())
        return random.choice([an.action for an in node.children.values() if an.value == top])


class _Step:
    """The fields of a model step the engine reads (one row of a batched ``model.step``)."""

    __slots__ = ("state", "observations", "rewards", "terminations", "truncations", "all_done")

    def __init__(self, state, observations, rew_row, flags, ids):
        self.state, self.observations = state, observations
        self.rewards = {i: float(rew_row[k]) for k, i in enumerate(ids)}
        self.terminations = {i: bool((int(flags) >> k) & 1) for k, i in enumerate(ids)}
        self.truncations = {i: False for i in ids}
        self.all_done = bool(int(flags) & 0x80)
