"""I-NTMCP in its general form -- any nesting level, any number of agents, any ``SearchPolicy`` objects -- on
the plugin path (reference: /root/reference/posggym_baselines/planning/intmcp.py:22-994).

The device planner (planning/intmcp.py) covers BASELINE.json's configuration: nesting level 1, two agents, random
search policies, two device trees.  Everything else is a hierarchy of policy trees, one per agent and level, whose
nodes are reached by walking *histories* and whose search policies are stateful Python objects; like
planning/host_search.py this keeps the trees and the plugin calls on the host and sends the model arithmetic to the
device:

* the initial beliefs of every tree of the hierarchy (``_initial_nested_update``, intmcp.py:216-276) are
  ``posg_belief_initial`` launches (one per initial history, the rejection loop of 231-243 inside the kernel);
* belief reinvigoration (843-892, belief.py:145-194) steps a round of rejection-sampling attempts per
  ``model.step`` launch, parents and joint actions chosen on the host in the sequential loop's order;
* single steps inside ``_simulate`` / ``_rollout`` go through the facade's scalar ``model.step``
  (``posg_model_step_host``); where every agent's search policy of a level is ``RandomSearchPolicy`` a leaf's whole
  rollout is ONE launch of the fused rollout kernel instead.

One deviation: the reference's PUCB selection reads ``self.action_space`` (intmcp.py:646), an attribute the class
never sets, so its I-NTMCP cannot search with ``action_selection="pucb"`` at all; here PUCB works (same formula as
``MCTS``).
"""
from __future__ import annotations

import ctypes as C
import math
import random
import time
from typing import Any, Dict, Optional

import numpy as np
import torch

from posggym_baselines_b200 import _lib
from posggym_baselines_b200.planning.config import MCTSConfig
from posggym_baselines_b200.planning.host_search import HPS, HostSearch, _ActionNode, _Belief, _MinMax, _ObsNode
from posggym_baselines_b200.planning.search_policy import RandomSearchPolicy, SearchPolicy
from posggym_baselines_b200.planning.utils import PlanningStatTracker


class _Draws:
    """Philox counters of the hierarchy's initial-belief launches (one key, disjoint counter ranges)."""

    def __init__(self, seed: int):
        self.seed, self.counter = seed & 0xFFFFFFFFFFFFFFFF, 0

    def take(self, n: int) -> int:
        c = self.counter
        self.counter += n
        return c


def _agent_history(joint_history, i):
    """((joint action, joint obs), ...) -> agent i's ((a, o), ...)."""
    return tuple((ja[i], jo[i]) for ja, jo in joint_history)


class HostINTMCP:
    """One policy tree of the hierarchy: the planner of ``agent_id`` at ``nesting_level``; its
    ``other_agent_policies`` are the trees one level down."""

    # node-level action selection is MCTS's (same formulas: intmcp.py:634-742 / mcts.py:494-600)
    _pucb, _ucb, _min_visit = HostSearch._pucb, HostSearch._ucb, HostSearch._min_visit
    _max_visit, _max_value = HostSearch._max_visit, HostSearch._max_value

    def __init__(self, model, agent_id: str, config: MCTSConfig, nesting_level: int,
                 other_agent_policies: Optional[Dict[str, "HostINTMCP"]], search_policies: Dict[str, SearchPolicy], *,
                 max_sims_per_level: Optional[int] = None, _draws: Optional[_Draws] = None):
        if not hasattr(model, "batched"):
            raise TypeError("model must be a posggym_baselines_b200 model facade (needs `.batched`)")
        other_agent_policies = other_agent_policies or {}
        self.model, self.agent_id, self.config, self.nesting_level = model, agent_id, config, nesting_level
        self.num_agents = len(model.possible_agents)
        assert nesting_level == 0 or len(other_agent_policies) == self.num_agents - 1
        assert agent_id not in other_agent_policies
        assert all(i in search_policies for i in model.possible_agents)
        self.other_agent_policies, self.search_policies = other_agent_policies, search_policies
        self.action_spaces = {i: list(range(model.action_spaces[i].n)) for i in model.possible_agents}
        self.action_space = self.action_spaces[agent_id]
        self._ids = list(model.possible_agents)
        self._ego = self._ids.index(agent_id)
        self._b = model.batched
        self._rng = random.Random(config.seed)
        if config.step_limit is not None:
            self.step_limit = config.step_limit
        elif model.spec is not None:
            self.step_limit = model.spec.max_episode_steps
        else:
            self.step_limit = float("inf")
        self._select = {"pucb": self._pucb, "ucb": self._ucb}.get(config.action_selection, self._min_visit)
        self._final = self._max_visit if config.action_selection == "pucb" else self._max_value
        self.max_sims_per_level = max_sims_per_level
        self._draws = _draws or _Draws(config.seed if config.seed is not None else random.getrandbits(48))
        # a rollout in which EVERY agent follows RandomSearchPolicy is MCTS._rollout with uniform-random policies:
        # the fused rollout kernel (posg_*_rollout), one launch instead of one scalar step per rollout step
        self._device_rollouts = all(type(p) is RandomSearchPolicy for p in search_policies.values())
        self.min_max = _MinMax(config.known_bounds)
        self.step_statistics: Dict[str, float] = {}
        self._reset_step_statistics()
        self.stat_tracker = PlanningStatTracker(self)
        self.reset()

    @classmethod
    def initialize(cls, model, ego_agent_id: str, config: MCTSConfig, nesting_level: int,
                   search_policies: Optional[Dict[int, Dict[str, SearchPolicy]]] = None, *,
                   max_sims_per_level: Optional[int] = None, _draws: Optional[_Draws] = None) -> "HostINTMCP":
        """The whole hierarchy below (and including) ``ego_agent_id`` at ``nesting_level`` (intmcp.py:949-994);
        ``search_policies[level][agent]``, None -> ``RandomSearchPolicy`` everywhere."""
        if search_policies is None:
            search_policies = {lvl: {i: RandomSearchPolicy(model, i) for i in model.possible_agents}
                               for lvl in range(nesting_level + 1)}
        draws = _draws or _Draws(config.seed if config.seed is not None else random.getrandbits(48))
        others = {}
        if nesting_level > 0:
            others = {i: cls.initialize(model, i, config, nesting_level - 1, search_policies,
                                        max_sims_per_level=max_sims_per_level, _draws=draws)
                      for i in model.possible_agents if i != ego_agent_id}
        return cls(model, ego_agent_id, config, nesting_level, others, search_policies[nesting_level],
                   max_sims_per_level=max_sims_per_level, _draws=draws)

    # ------------------------------------------------------------------ step / reset
    def step(self, obs):
        root = self.traverse(self.history)
        assert self.step_limit is None or root.t <= self.step_limit
        if root.is_absorbing:
            for k in self.step_statistics:
                self.step_statistics[k] = np.nan
            return self._last_action
        self._reset_step_statistics()
        self.update(self._last_action, obs)
        self._last_action = self.get_action()
        self._step_num += 1
        self.step_statistics.update(self._collect_nested_statistics())
        self.step_statistics["mem_usage"] = self._mem_usage()
        self.stat_tracker.step()
        return self._last_action

    @staticmethod
    def _mem_usage() -> float:
        try:
            import psutil

            return psutil.Process().memory_info().rss / 1024**2
        except Exception:
            return 0.0

    def _collect_nested_statistics(self) -> Dict[str, float]:
        """Only what every level collects is summed over the hierarchy (intmcp.py:138-152)."""
        stats = {"reinvigoration_time": self.step_statistics["reinvigoration_time"]}
        for pi in self.other_agent_policies.values():
            stats["reinvigoration_time"] += pi._collect_nested_statistics()["reinvigoration_time"]
        return stats

    def reset(self):
        self.stat_tracker.reset_episode()
        self._step_num = 0
        self.min_max = _MinMax(self.config.known_bounds)
        self._reset_step_statistics()
        self.root = _ObsNode(None, None, 0, _Belief(self._rng), {None: 1.0},
                             self.search_policies[self.agent_id].get_initial_state())
        self._last_action = None
        self.history: tuple = ()
        for pi in self.other_agent_policies.values():
            pi.reset()

    def _reset_step_statistics(self):
        self.step_statistics = {
            "search_time": 0.0, "update_time": 0.0, "reinvigoration_time": 0.0, "evaluation_time": 0.0,
            "policy_calls": 0, "inference_time": 0.0, "search_depth": 0, "num_sims": 0, "mem_usage": 0,
            "min_value": self.min_max.minimum, "max_value": self.min_max.maximum}

    # ------------------------------------------------------------------ update
    def update(self, action, obs):
        root = self.traverse(self.history)
        if root.is_absorbing:
            return
        start = time.time()
        if root.t == 0:
            self.history = ((None, obs),)
            self._initial_nested_update({self.history: 1.0})
        else:
            self.history = self.history + ((action, obs),)
            self._nested_update({self.history: 1.0}, len(self.history))
        self.step_statistics["update_time"] = time.time() - start

    def _initial_nested_update(self, history_dist: Dict[tuple, float]):
        cfg, b, ids = self.config, self._b, self._ids
        init_actions = dict.fromkeys(ids)
        for hist, h_prob in history_dist.items():
            node = self.traverse(hist)
            init_obs = hist[-1][1]
            need = max(0, math.ceil(h_prob * (cfg.num_particles + cfg.extra_particles)))   # `while size < h_prob * N`
            belief = _Belief(self._rng)
            if need:
                states = self._device_initial_belief(init_obs, need)
                joint_obs_all, unpacked = b.unpack_obs(b.observe(states)), b.unpack_state(states)
                for k in range(need):
                    joint_obs = {i: joint_obs_all[k][n] for n, i in enumerate(ids)}
                    joint_obs[self.agent_id] = init_obs
                    pstate = {j: self.search_policies[j].get_initial_state() for j in ids if j != self.agent_id}
                    pstate = self._others_next(init_actions, joint_obs, pstate)
                    belief.add_particle(HPS(unpacked[k], ((dict(init_actions), joint_obs),), pstate, 1))
            node.belief = belief
        if self.nesting_level > 0:
            nested = self.get_nested_history_dist(history_dist)
            for i, pi in self.other_agent_policies.items():
                pi._initial_nested_update(nested[i])

    def _device_initial_belief(self, init_obs, need: int) -> torch.Tensor:
        """``need`` packed initial states whose first observation for this tree's agent is ``init_obs``: the
        rejection loop of intmcp.py:231-243 in one launch."""
        b = self._b
        sw = C.c_int()
        _lib.check(_lib.lib().posg_model_info(b._handle, C.byref(C.c_int()), C.byref(C.c_int()), C.byref(sw)))
        out = torch.empty((need, sw.value), dtype=torch.int32, device=b.device)
        res = np.zeros(3, np.int64)
        key = np.ascontiguousarray(b.pack_obs(init_obs), dtype=np.uint64)
        max_attempts = 1 << 24
        with torch.cuda.device(b.device):
            _lib.check(_lib.lib().posg_belief_initial(b._handle, self._ego, _lib.ptr(key), need, max_attempts,
                                                      self._draws.seed, self._draws.take(max_attempts), _lib.ptr(out),
                                                      _lib.ptr(res), _lib.stream_ptr()))
        if res[0] < need:
            raise RuntimeError(f"initial belief: only {res[0]} of {need} particles matched the observation")
        return out

    def _nested_update(self, history_dist: Dict[tuple, float], current_t: int):
        # the reference first walks the tree dropping references to dead histories and clearing stale beliefs
        # (intmcp.py:279-333); that only frees memory and is left to the garbage collector here
        target = self.config.num_particles + self.config.extra_particles
        for hist, h_prob in history_dist.items():
            node = self.traverse(hist)
            if node.is_absorbing:
                continue
            action, obs = hist[-1]
            self._reinvigorate(node, action, obs, math.ceil(h_prob * target))
        if self.nesting_level > 0:
            nested = self.get_nested_history_dist(history_dist)
            for i, pi in self.other_agent_policies.items():
                pi._nested_update(nested[i], current_t)

    def get_nested_history_dist(self, histories: Dict[tuple, float]) -> Dict[str, Dict[tuple, float]]:
        """Distribution over each other agent's histories implied by the beliefs at ``histories`` (intmcp.py:336-362)."""
        nested: Dict[str, Dict[tuple, float]] = {}
        for hist, h_prob in histories.items():
            belief = self.traverse(hist).belief
            size = belief.size()
            for i in self.other_agent_policies:
                dist = nested.setdefault(i, {})
                counts: Dict[tuple, int] = {}
                for hps in belief.particles:
                    h = _agent_history(hps.history, i)
                    counts[h] = counts.get(h, 0) + 1
                for h, c in counts.items():
                    dist[h] = dist.get(h, 0) + h_prob * (c / size)
        return nested

    # ------------------------------------------------------------------ search
    def get_action(self):
        """Top-level tree only: an equal share of the budget per level, lowest level first (intmcp.py:368-408)."""
        root = self.traverse(self.history)
        if root.is_absorbing:
            return self.action_space[0]
        start = time.time()
        per_level = self.config.search_time_limit / (self.nesting_level + 1)
        n_sims = 0
        for level in range(self.nesting_level + 1):
            t0, k = time.time(), 0
            while (time.time() - t0 < per_level) if self.max_sims_per_level is None else (k < self.max_sims_per_level):
                self._nested_sim(self.history, level, True)
                k += 1
            n_sims += k
        self.step_statistics.update({"search_time": time.time() - start, "num_sims": n_sims,
                                     "min_value": self.min_max.minimum, "max_value": self.min_max.maximum})
        return self._final(root)

    def _nested_sim(self, history: tuple, search_level: int, top_level: bool = False):
        root = self.traverse(history)
        if not root.children:
            for a in self.action_space:
                root.add_child(a)
        size = root.belief.size()
        if size == 0 or (top_level and size < self.config.extra_particles):       # depleted node
            self._reinvigorate(root, root.parent.action, root.obs, self.config.extra_particles)
        hps = root.belief.sample()
        if self.nesting_level > search_level:
            for i, pi in self.other_agent_policies.items():
                pi._nested_sim(_agent_history(hps.history, i), search_level, False)
        else:
            _, depth = self._simulate(hps, root, 0)
            root.visits += 1
            self.step_statistics["search_depth"] = max(self.step_statistics["search_depth"], depth)

    def _simulate(self, hps: HPS, node: _ObsNode, depth: int):
        cfg = self.config
        if depth > cfg.depth_limit or (self.step_limit is not None and node.t + depth > self.step_limit):
            return 0, depth
        if len(node.children) < len(self.action_space):
            # a higher tree's queries may have added some of the action nodes already
            for a in self.action_space:
                if a not in node.children:
                    node.add_child(a)
            return self._evaluate(hps, depth, node), depth
        ego_action = self._select(node)
        joint_action = self._get_joint_action(hps, ego_action)
        ts = self.model.step(hps.state, joint_action)
        ego_obs, ret = ts.observations[self.agent_id], ts.rewards[self.agent_id]
        done = ts.terminations[self.agent_id] or ts.truncations[self.agent_id] or ts.all_done
        nxt = self._advance(hps, joint_action, ts.state, ts.observations)
        anode = node.children[ego_action]
        child = anode.children.get(ego_obs)
        if child is not None:
            child.visits += 1
            if cfg.action_selection == "pucb":
                pi = self.search_policies[self.agent_id].get_pi(child.search_policy_state)
                for a, pr in pi.items():
                    child.action_probs[a] += (pr - child.action_probs[a]) / child.visits
        else:
            child = self._add_obs_node(anode, ego_obs, 1)
        child.is_absorbing = done
        child.belief.add_particle(nxt)
        max_depth = depth
        if not done:
            future, max_depth = self._simulate(nxt, child, depth + 1)
            ret += cfg.discount * future
        anode.update(ret)
        self.min_max.update(anode.value)
        return ret, max_depth

    def _evaluate(self, hps: HPS, depth: int, node: _ObsNode) -> float:
        t0, cfg = time.time(), self.config

        def rollout():
            if self._device_rollouts:
                return self._device_rollout(hps, depth)
            states = dict(hps.policy_state)
            states[self.agent_id] = node.search_policy_state
            return self._rollout(hps, depth, states)

        if cfg.truncated:
            try:
                v = self.search_policies[self.agent_id].get_value(node.search_policy_state)
            except NotImplementedError:
                if not cfg.use_rollout_if_no_value:
                    raise
                v = rollout()
        else:
            v = rollout()
        self.step_statistics["evaluation_time"] += time.time() - t0
        return v

    def _rollout(self, hps: HPS, depth: int, pstates: Dict[str, Any]) -> float:
        """Every agent acts by its search policy of this level (intmcp.py:547-593)."""
        cfg, ret, k, state, t = self.config, 0, 0, hps.state, hps.t
        while depth <= cfg.depth_limit and t <= self.step_limit:
            joint_action = {i: self.search_policies[i].sample_action(pstates[i]) for i in self._ids}
            ts = self.model.step(state, joint_action)
            ret += cfg.discount**k * ts.rewards[self.agent_id]
            if ts.terminations[self.agent_id] or ts.truncations[self.agent_id] or ts.all_done:
                break
            # the reference also advances the particle's own copy of the other agents' states here
            # (intmcp.py:574-577) -- the same calls on the same policies with the same results as the line below,
            # so they are counted, not repeated
            self.step_statistics["policy_calls"] += len(self._ids) - 1
            pstates = {i: self._policy_next(joint_action[i], ts.observations[i], self.search_policies[i], pstates[i])
                       for i in self._ids}
            state, t = ts.state, t + 1
            depth += 1
            k += 1
        return ret

    def _device_rollout(self, hps: HPS, depth: int) -> float:
        """``_rollout`` when all agents act uniformly at random: the loop's bounds (``depth <= depth_limit and
        t <= step_limit``), discounting and stop rule inside the fused rollout kernel, actions from a Philox
        sub-sequence of this rollout's own."""
        cfg, b = self.config, self._b
        steps = min(cfg.depth_limit - depth + 1, self.step_limit - hps.t + 1, 255)
        if steps <= 0:
            return 0
        st = torch.from_numpy(b.pack_state([hps.state])).to(b.device)
        ret, done, _ = b.rollout(st, self._ego, int(steps), cfg.discount, seed=self._draws.seed ^ 0x726F6C6C,
                                 particle_offset=self._draws.take(1))
        k = int(done.cpu()[0])
        # the reference advances every agent's (stateless) search-policy state after each step that does not end the
        # rollout: 2 * agents - 1 counted calls per such step (intmcp.py:574-588)
        self.step_statistics["policy_calls"] += (2 * len(self._ids) - 1) * max(0, k - 1 if k < steps else k)
        return float(ret.cpu()[0])

    # ------------------------------------------------------------------ policies and joint actions
    def _policy_next(self, action, obs, policy, state):
        t0 = time.time()
        nxt = policy.get_next_state(action, obs, state)
        self.step_statistics["inference_time"] += time.time() - t0
        self.step_statistics["policy_calls"] += 1
        return nxt

    def _others_next(self, joint_action, joint_obs, pstate) -> Dict[str, Any]:
        """Next search-policy state of every other agent; the ego's slot stays None (intmcp.py:610-628)."""
        return {i: (None if i == self.agent_id or i not in joint_action else
                    self._policy_next(joint_action[i], joint_obs[i], self.search_policies[i], pstate[i]))
                for i in self._ids}

    def _advance(self, hps: HPS, joint_action, state, joint_obs) -> HPS:
        return HPS(state, hps.history + ((dict(joint_action), dict(joint_obs)),),
                   self._others_next(joint_action, joint_obs, hps.policy_state), hps.t + 1)

    def _get_joint_action(self, hps: HPS, ego_action) -> Dict[str, int]:
        """Other agents act from their trees one level down; random at level 0 (intmcp.py:744-761, 894-912)."""
        out = {}
        for i in self._ids:
            if i == self.agent_id:
                out[i] = ego_action
            elif self.nesting_level == 0 or self.config.state_belief_only:
                out[i] = self._rng.choice(self.action_spaces[i])
            else:
                out[i] = self.other_agent_policies[i].sample_action(_agent_history(hps.history, i),
                                                                    self.search_policies[i], hps.policy_state[i])
        return out

    def sample_action(self, history: tuple, search_policy: SearchPolicy, search_policy_state):
        """This tree's policy at ``history``: softmax of the action visits over sqrt(node visits); the caller's search
        policy where the tree has nothing yet (intmcp.py:763-791)."""
        node = self.traverse(history)
        if node.visits == 0 or not node.children:
            return search_policy.sample_action(search_policy_state)
        root_n = math.sqrt(node.visits)
        kids = list(node.children.values())
        w = [math.exp(an.visits / root_n) for an in kids]
        z = sum(w)
        return random.choices(kids, weights=[x / z for x in w])[0].action

    # ------------------------------------------------------------------ tree
    def traverse(self, history: tuple) -> _ObsNode:
        """The node ``history`` leads to, created on the way if missing (intmcp.py:797-809)."""
        node = self.root
        for a, o in history:
            anode = node.children.get(a)
            if anode is None:
                anode = node.add_child(a)
            nxt = anode.children.get(o)
            node = nxt if nxt is not None else self._add_obs_node(anode, o, 0)
        return node

    def _add_obs_node(self, parent: _ActionNode, obs, init_visits: int = 0) -> _ObsNode:
        policy = self.search_policies[self.agent_id]
        sp_state = self._policy_next(parent.action, obs, policy, parent.parent.search_policy_state)
        node = _ObsNode(parent, obs, parent.parent.t + 1, _Belief(self._rng), policy.get_pi(sp_state), sp_state, init_visits)
        parent.children[obs] = node
        return node

    # ------------------------------------------------------------------ reinvigoration
    def _reinvigorate(self, node: _ObsNode, action, obs, target_node_size: Optional[int] = None):
        """``BeliefRejectionSampler._rejection_sample`` with ``use_rejected_samples=True`` (belief.py:145-194), a round
        of attempts per ``model.step`` launch."""
        t0 = time.time()
        cfg, b, ids = self.config, self._b, self._ids
        need = (cfg.num_particles + cfg.extra_particles if target_node_size is None else target_node_size) - node.belief.size()
        if need > 0:
            parent = node.parent.parent.belief
            assert parent is not None
            limit = cfg.reinvigoration_sample_limit_factor * need
            want = b.pack_obs(obs)
            accepted, rejected, attempts = [], [], 0
            while len(accepted) < need and attempts < limit:
                n = int(min(math.ceil(limit - attempts), max(32, 2 * (need - len(accepted)))))
                parents = [parent.sample() for _ in range(n)]
                actions = [self._get_joint_action(h, action) for h in parents]
                act = np.array([[int(a[i]) for i in ids] for a in actions], dtype=np.uint8)
                nst, nob, _, _ = b.step_host(b.pack_state([h.state for h in parents]), act, self.model.host_step_noise(n))
                keys = np.ascontiguousarray(nob.reshape(n, len(ids), -1)[:, self._ego, :]).view(np.uint64)
                match = (keys == want.reshape(1, -1)).all(axis=1)
                nstates, nobs = b.unpack_state(nst), b.unpack_obs(nob)
                for k in range(n):
                    if len(accepted) >= need or attempts >= limit:      # the sequential loop stops here
                        break
                    attempts += 1
                    joint_obs = {i: nobs[k][m] for m, i in enumerate(ids)}
                    nxt = self._advance(parents[k], actions[k], nstates[k], joint_obs)
                    (accepted if match[k] else rejected).append(nxt)
            if len(accepted) < need:
                accepted.extend(rejected[: need - len(accepted)])
            for hps in accepted:
                node.belief.add_particle(hps)
        self.step_statistics["reinvigoration_time"] += time.time() - t0

    # ------------------------------------------------------------------ misc
    def close(self):
        for policy in self.search_policies.values():
            policy.close()
        for policy in self.other_agent_policies.values():
            policy.close()

    def __str__(self):
        return "INTMCP"

