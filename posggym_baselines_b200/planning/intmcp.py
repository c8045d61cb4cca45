"""``INTMCP`` (Interactive Nested Tree Monte-Carlo Planning) on the GPU, nesting level 1,
two agents -- BASELINE.json configs[3] (reference: /root/reference/posggym_baselines/
planning/intmcp.py:22-994; smoke scenario tests/planning/test_intmcp.py:34-48).

Same surface as the reference: ``INTMCP.initialize(model, ego_agent_id, config,
nesting_level, search_policies)``, ``reset()``, ``step(obs)``, ``update(action, obs)``,
``get_action()``, ``close()``, ``step_statistics``, ``stat_tracker``.

What runs where:

* two device trees (csrc/search.cuh): the ego agent's level-1 tree and the other agent's
  level-0 tree;
* a level-1 particle is ``(packed state, node of the level-0 tree the other agent's
  history leads to)`` -- the reference's ``(state, JointHistory)`` particle, because
  ``INTMCP.traverse(history)`` is a pure function of the history (intmcp.py:797-809);
* ``get_action`` (intmcp.py:368-408): half of the budget grows the level-0 tree from the
  nodes the root belief's particles point to (``posg_search_multi``), the other half
  searches the level-1 tree with the other agent acting by softmax over the level-0
  node's visit counts (``posg_search_nested``; ``sample_action`` 763-791);
* ``update`` (198-304): re-root + reinvigoration of the level-1 belief with the same
  nested action model, then the nested history distribution (336-362) and the level-0
  beliefs of exactly those histories (gathered from the level-0 particle log, topped up
  by rejection sampling from the parent history's belief).

Only ``RandomSearchPolicy`` search policies are supported (``search_policies=None``),
``state_belief_only`` must be False (as the reference requires for nesting), nesting level
1, two agents.  Pruning of dead histories (intmcp.py:305-333) only frees memory in the
reference and is not needed in a fixed arena.
"""
from __future__ import annotations

import ctypes as C
import math
import random
import time
from typing import Dict, Optional

import numpy as np
import torch

from posggym_baselines_b200 import _lib
from posggym_baselines_b200.planning.config import MCTSConfig
from posggym_baselines_b200.planning.search_policy import RandomSearchPolicy, SearchPolicy
from posggym_baselines_b200.planning.utils import PlanningStatTracker

_SELECTION = {"ucb": 0, "pucb": 1, "uniform": 2}


class INTMCP:
    def __init__(self, model, agent_id: str, config: MCTSConfig, nesting_level: int,
                 other_agent_policies: Optional[Dict[str, "INTMCP"]], search_policies: Dict[str, SearchPolicy], *,
                 max_sims_per_level: Optional[int] = None, wave_size: Optional[int] = None, max_wave: int = 262144,
                 node_capacity: int = 1 << 25, log_capacity: int = 1 << 27, belief_cap: int = 4096,
                 max_histories: int = 64, history_belief_cap: int = 256):
        if nesting_level != 1 or len(model.possible_agents) != 2:
            raise NotImplementedError("the device I-NTMCP is nesting_level=1 with two agents; INTMCP.initialize(...) "
                                      "builds any other hierarchy on the plugin path (planning/host_intmcp.py)")
        if not all(isinstance(p, RandomSearchPolicy) for p in search_policies.values()):
            raise NotImplementedError("only RandomSearchPolicy search policies run in-kernel; INTMCP.initialize(...) "
                                      "routes plugin search policies to planning/host_intmcp.py")
        if config.state_belief_only:
            raise ValueError("I-NTMCP needs history beliefs: state_belief_only must be False")
        if not hasattr(model, "batched"):
            raise TypeError("model must be a posggym_baselines_b200 model facade (needs `.batched`)")
        self.model, self.agent_id, self.config, self.nesting_level = model, agent_id, config, nesting_level
        self.search_policies = search_policies
        self.other_agent_policies = other_agent_policies or {}
        self.num_agents = 2
        self._b = model.batched
        ids = list(model.possible_agents)
        self._ego = ids.index(agent_id)
        self._other = 1 - self._ego
        self.action_spaces = {i: list(range(model.action_spaces[i].n)) for i in ids}
        if config.step_limit is not None:
            self.step_limit = config.step_limit
        elif model.spec is not None:
            self.step_limit = model.spec.max_episode_steps
        else:
            self.step_limit = 250
        self._seed = config.seed if config.seed is not None else random.getrandbits(48)
        self.wave_size, self.max_wave = wave_size, max_wave
        self._cap_nodes, self._cap_log = int(node_capacity), int(log_capacity)
        self.max_sims_per_level = max_sims_per_level
        # the reference's beliefs hold tens of particles; a search that runs millions of
        # simulations leaves millions at a child, so beliefs are capped (a subset of a particle
        # set is a particle set of the same distribution) and only the `max_histories` most
        # probable histories of the other agent keep a level-0 belief
        self.belief_cap, self.max_histories, self.history_belief_cap = belief_cap, max_histories, history_belief_cap

        lib = _lib.lib()
        self._nested = C.c_void_p()
        _lib.check(lib.posg_nested_create(self._b._handle, C.byref(self._nested)))
        na, ow, sw = C.c_int(), C.c_int(), C.c_int()
        _lib.check(lib.posg_model_info(self._b._handle, C.byref(na), C.byref(ow), C.byref(sw)))
        self._A, self._OW, self._SW = na.value, ow.value, sw.value
        self._arena1: Optional[torch.Tensor] = None
        self._arena0: Optional[torch.Tensor] = None

        def params(ego):
            p = _lib.SearchParams()
            kb = config.known_bounds
            p.discount, p.c = config.discount, config.c
            p.known_min, p.known_max = (kb.min, kb.max) if kb else (1.0, 0.0)
            p.pucb_exploration_fraction = config.pucb_exploration_fraction
            p.depth_limit = config.depth_limit
            p.step_limit = int(min(self.step_limit, 250))
            p.selection = _SELECTION[config.action_selection]
            p.ego_agent = ego
            p.intmcp_depth_rule = 1
            p.seed = (self._seed + 7919 * ego) & 0xFFFFFFFFFFFFFFFF
            return p

        self._params1, self._params0 = params(self._ego), params(self._other)
        self.step_statistics: Dict[str, float] = {}
        self.stat_tracker = PlanningStatTracker(self)
        self.reset()

    @classmethod
    def initialize(cls, model, ego_agent_id: str, config: MCTSConfig, nesting_level: int,
                   search_policies: Optional[Dict[int, Dict[str, SearchPolicy]]] = None, **b200_options) -> "INTMCP":
        """Reference signature (intmcp.py:949-994); ``search_policies=None`` -> RandomSearchPolicy.  Nesting level 1
        with two agents and random search policies is the device planner (this class); any other hierarchy -- deeper
        nesting, more agents, plugin search policies -- is built by planning/host_intmcp.py (host trees, model on the
        device), which has the same surface."""
        if search_policies is None:
            search_policies = {lvl: {i: RandomSearchPolicy(model, i) for i in model.possible_agents}
                               for lvl in range(nesting_level + 1)}
        in_kernel = (nesting_level == 1 and len(model.possible_agents) == 2 and all(
            type(p) is RandomSearchPolicy for lvl in range(nesting_level + 1) for p in search_policies[lvl].values()))
        if not in_kernel:
            from posggym_baselines_b200.planning.host_intmcp import HostINTMCP

            return HostINTMCP.initialize(model, ego_agent_id, config, nesting_level, search_policies,
                                         max_sims_per_level=b200_options.get("max_sims_per_level"))
        return cls(model, ego_agent_id, config, nesting_level, None, search_policies[nesting_level], **b200_options)

    # ------------------------------------------------------------------ plumbing
    def _dev(self):
        return torch.cuda.device(self._b.device)

    def _stream(self):
        return _lib.stream_ptr()

    def _ensure_arenas(self):
        lib = _lib.lib()
        if self._arena1 is None:
            n1 = _lib.check(int(lib.posg_tree_arena_bytes(self._nested, self._cap_nodes, self._cap_log)))
            n0 = _lib.check(int(lib.posg_tree_arena_bytes(self._b._handle, self._cap_nodes, self._cap_log)))
            self._arena1 = torch.empty(n1, dtype=torch.uint8, device=self._b.device)
            self._arena1_spare = torch.empty(n1, dtype=torch.uint8, device=self._b.device)
            # the spare must be clean once: posg_tree_reroot only re-initialises the slots its previous life used
            with self._dev():
                _lib.check(lib.posg_tree_reset(self._nested, _lib.ptr(self._arena1_spare), self._cap_nodes,
                                               self._cap_log, 0, 1.0, 0.0, self._stream()))
            self._scratch = torch.empty(3 * self._cap_nodes, dtype=torch.int32, device=self._b.device)
            self._arena0 = torch.empty(n0, dtype=torch.uint8, device=self._b.device)

    def _stats(self, level: int, node: int):
        out = np.zeros(3 * self._A + 10, np.float64)
        handle, arena = (self._nested, self._arena1) if level == 1 else (self._b._handle, self._arena0)
        with self._dev():
            _lib.check(_lib.lib().posg_tree_node_stats(handle, _lib.ptr(arena), self._cap_nodes, self._cap_log, node,
                                                       _lib.ptr(out), self._stream()))
        A, e = self._A, out[3 * self._A:]
        return {"done": out[A:2 * A].copy(), "total": out[2 * A:3 * A].copy(), "node_visits": e[0],
                "node_count": e[1], "log_count": e[2], "max_depth": e[3], "min_value": e[4], "max_value": e[5],
                "sims_done": e[7], "flags": int(e[8])}

    def _next_draw(self, n):
        d = self._draw
        self._draw += int(n)
        return d

    # ------------------------------------------------------------------ reset / step
    def reset(self):
        self.stat_tracker.reset_episode()
        self._step_num = 0
        self._root1, self._root_t, self._absorbing = 0, 0, False
        self._P1: Optional[torch.Tensor] = None
        self._t0_beliefs: Dict[int, torch.Tensor] = {}
        self._last_action = None
        self._sim0 = self._sim1 = self._draw = 0
        self._min_value, self._max_value = float("inf"), -float("inf")
        self._reset_step_statistics()

    def _reset_step_statistics(self):
        self.step_statistics = {
            "search_time": 0.0, "update_time": 0.0, "reinvigoration_time": 0.0, "evaluation_time": 0.0,
            "policy_calls": 0, "inference_time": 0.0, "search_depth": 0, "num_sims": 0, "mem_usage": 0,
            "min_value": self._min_value, "max_value": self._max_value}

    def step(self, obs):
        if self._absorbing:
            for k in self.step_statistics:
                self.step_statistics[k] = np.nan
            return self._last_action
        self._reset_step_statistics()
        self.update(self._last_action, obs)
        self._last_action = self.get_action()
        self._step_num += 1
        self.stat_tracker.step()
        return self._last_action

    # ------------------------------------------------------------------ update
    def update(self, action, obs):
        if self._absorbing:
            return
        start = time.time()
        if self._root_t == 0:
            self._initial_update(obs)
        else:
            self._nested_update(action, obs)
        torch.cuda.synchronize(self._b.device)
        self.step_statistics["update_time"] = time.time() - start

    def _initial_belief(self, agent: int, key: np.ndarray, need: int) -> torch.Tensor:
        out = torch.empty((need, self._SW), dtype=torch.int32, device=self._b.device)
        res = np.zeros(3, np.int64)
        max_attempts = 1 << 24
        with self._dev():
            _lib.check(_lib.lib().posg_belief_initial(
                self._b._handle, agent, _lib.ptr(key), need, max_attempts, self._seed & 0xFFFFFFFFFFFFFFFF,
                self._next_draw(max_attempts), _lib.ptr(out), _lib.ptr(res), self._stream()))
        if res[0] < need:
            raise RuntimeError(f"initial belief: only {res[0]} of {need} particles matched the observation")
        return out

    def _initial_update(self, init_obs):
        cfg, b, lib = self.config, self._b, _lib.lib()
        self._ensure_arenas()
        kb = cfg.known_bounds
        kmin, kmax = (kb.min, kb.max) if kb else (1.0, 0.0)
        n_total = cfg.num_particles + cfg.extra_particles
        with self._dev():
            _lib.check(lib.posg_tree_reset(self._nested, _lib.ptr(self._arena1), self._cap_nodes, self._cap_log, 1,
                                           kmin, kmax, self._stream()))
            _lib.check(lib.posg_tree_reset(b._handle, _lib.ptr(self._arena0), self._cap_nodes, self._cap_log, 0,
                                           kmin, kmax, self._stream()))
        key = np.ascontiguousarray(b.pack_obs(init_obs), dtype=np.uint64)
        states = self._initial_belief(self._ego, key, n_total)                  # [N, SW]
        other_obs = b.observe(states).reshape(n_total, self.num_agents, -1)[:, self._other, :].contiguous()
        zeros = torch.zeros(n_total, dtype=torch.int32, device=b.device)
        n0 = torch.empty(n_total, dtype=torch.int32, device=b.device)
        with self._dev():
            _lib.check(lib.posg_tree_advance_batch(b._handle, _lib.ptr(self._arena0), self._cap_nodes, self._cap_log,
                                                   n_total, _lib.ptr(zeros), _lib.ptr(zeros), _lib.ptr(other_obs),
                                                   _lib.ptr(n0), self._stream()))
        self._P1 = torch.cat([states, n0.view(-1, 1)], dim=1).contiguous()
        # _initial_nested_update of the other agent's tree (intmcp.py:273-276): a belief per
        # initial history of the other agent, sized h_prob * (num_particles + extra_particles)
        self._t0_beliefs = {}
        n0_h, obs_h = n0.cpu().numpy(), other_obs.cpu().numpy().astype(np.int64).view(np.uint64)
        for u in np.unique(n0_h):
            idx = int(np.nonzero(n0_h == u)[0][0])
            need = int((n0_h == u).sum())
            self._t0_beliefs[int(u)] = self._initial_belief(self._other, np.ascontiguousarray(obs_h[idx]), need)
        self._root1, self._root_t, self._absorbing = 0, 1, False
        self._nodes1 = 1

    def _nested_update(self, action, obs):
        cfg, b, lib = self.config, self._b, _lib.lib()
        n_total = cfg.num_particles + cfg.extra_particles
        key = np.ascontiguousarray(b.pack_obs(obs), dtype=np.uint64)
        out4 = np.zeros(4, np.int32)
        with self._dev():
            _lib.check(lib.posg_tree_advance(self._nested, _lib.ptr(self._arena1), self._cap_nodes, self._cap_log,
                                             self._root1, int(action), _lib.ptr(key), 1, _lib.ptr(out4), self._stream()))
        child, absorbing = int(out4[0]), bool(int(out4[2]) & 2)
        if child < 0:
            raise RuntimeError("level-1 tree arena is full; raise node_capacity")
        self._root1, self._root_t, self._absorbing = child, self._root_t + 1, absorbing
        if absorbing:
            return
        t0 = time.time()
        sw1 = self._SW + 1
        buf = torch.empty((self.belief_cap + n_total, sw1), dtype=torch.int32, device=b.device)
        cnt = np.zeros(1, np.int64)
        with self._dev():
            _lib.check(lib.posg_tree_gather_belief(self._nested, _lib.ptr(self._arena1), self._cap_nodes, self._cap_log,
                                                   child, _lib.ptr(buf), buf.shape[0] - n_total, _lib.ptr(cnt),
                                                   self._stream()))
        have = min(int(cnt[0]), self.belief_cap)
        need = n_total - have
        if need > 0:
            limit = int(cfg.reinvigoration_sample_limit_factor * need)
            rej = torch.empty((need, sw1), dtype=torch.int32, device=b.device)
            res = np.zeros(3, np.int64)
            with self._dev():
                _lib.check(lib.posg_belief_reinvigorate_nested(
                    self._nested, _lib.ptr(self._arena0), self._cap_nodes, self._cap_log, self._other,
                    _lib.ptr(self._P1), self._P1.shape[0], self._ego, int(action), _lib.ptr(key), need, limit,
                    (self._seed ^ 0x5DEECE66D) & 0xFFFFFFFFFFFFFFFF, self._next_draw(limit), _lib.ptr(buf[have:]),
                    _lib.ptr(rej), _lib.ptr(res), self._stream()))
            have += int(res[2])
        self._P1 = buf[:have]
        # keep the level-1 subtree under the new root, drop the rest (intmcp.py:305-333 prunes the
        # dead histories); the level-0 tree is append-only because level-1 particles index it
        rr = np.zeros(2, np.int64)
        with self._dev():
            _lib.check(lib.posg_tree_reroot(self._nested, _lib.ptr(self._arena1), _lib.ptr(self._arena1_spare),
                                            self._cap_nodes, self._cap_log, child, _lib.ptr(self._scratch),
                                            _lib.ptr(rr), self._stream()))
        self._arena1, self._arena1_spare = self._arena1_spare, self._arena1
        self._root1 = 0
        self._nodes1 = int(rr[0])
        # nested history distribution (intmcp.py:336-362) and the other tree's beliefs (278-304)
        n0 = self._P1[:, self._SW].contiguous()
        n0_h = n0.cpu().numpy()
        uniq, counts = np.unique(n0_h[n0_h >= 0], return_counts=True)
        if len(uniq) > self.max_histories:
            keep = np.argsort(-counts, kind="stable")[: self.max_histories]
            uniq, counts = uniq[keep], counts[keep]
        new_beliefs: Dict[int, torch.Tensor] = {}
        if len(uniq):
            nodes = torch.from_numpy(uniq.astype(np.int32)).to(b.device)
            par = torch.empty(len(uniq), dtype=torch.int32, device=b.device)
            act = torch.empty_like(par)
            tt = torch.empty_like(par)
            ob = torch.empty((len(uniq), self._OW), dtype=torch.int64, device=b.device)
            with self._dev():
                _lib.check(lib.posg_tree_node_info(b._handle, _lib.ptr(self._arena0), self._cap_nodes, self._cap_log,
                                                   len(uniq), _lib.ptr(nodes), _lib.ptr(par), _lib.ptr(act),
                                                   _lib.ptr(ob), _lib.ptr(tt), self._stream()))
            par_h, act_h = par.cpu().numpy(), act.cpu().numpy()
            ob_h = ob.cpu().numpy().view(np.uint64)
            n_p1 = len(n0_h)
            hcap = self.history_belief_cap
            for j, u in enumerate(uniq):
                target = math.ceil(counts[j] / n_p1 * n_total)
                bbuf = torch.empty((hcap + target, self._SW), dtype=torch.int32, device=b.device)
                c0 = np.zeros(1, np.int64)
                with self._dev():
                    _lib.check(lib.posg_tree_gather_belief(b._handle, _lib.ptr(self._arena0), self._cap_nodes,
                                                           self._cap_log, int(u), _lib.ptr(bbuf),
                                                           bbuf.shape[0] - target, _lib.ptr(c0), self._stream()))
                h = min(int(c0[0]), hcap)
                miss = target - h
                parent_belief = self._t0_beliefs.get(int(par_h[j]))
                if miss > 0 and parent_belief is not None and parent_belief.shape[0] > 0:
                    lim = int(cfg.reinvigoration_sample_limit_factor * miss)
                    rej0 = torch.empty((miss, self._SW), dtype=torch.int32, device=b.device)
                    r0 = np.zeros(3, np.int64)
                    with self._dev():
                        _lib.check(lib.posg_belief_reinvigorate(
                            b._handle, _lib.ptr(parent_belief), parent_belief.shape[0], self._other, int(act_h[j]),
                            _lib.ptr(np.ascontiguousarray(ob_h[j])), miss, lim,
                            (self._seed ^ 0xB5297A4D) & 0xFFFFFFFFFFFFFFFF, self._next_draw(lim),
                            _lib.ptr(bbuf[h:]), _lib.ptr(rej0), _lib.ptr(r0), self._stream()))
                    h += int(r0[2])
                elif miss > 0:
                    # no parent belief to reinvigorate from: the level-1 particles that carry this
                    # history are themselves consistent with it
                    own = self._P1[n0 == int(u)][:, :self._SW]
                    take = min(miss, own.shape[0])
                    bbuf[h:h + take] = own[:take]
                    h += take
                new_beliefs[int(u)] = bbuf[:h]
        self._t0_beliefs = new_beliefs
        self.step_statistics["reinvigoration_time"] += time.time() - t0

    # ------------------------------------------------------------------ search
    def _budget(self) -> int:
        if self.max_sims_per_level is not None:
            return int(self.max_sims_per_level)
        # every simulation adds at most one node to its tree; spread the arena over the episode
        return max(1024, self._cap_nodes // (int(self.step_limit) + 2))

    def get_action(self):
        if self._absorbing:
            return self.action_spaces[self.agent_id][0]
        cfg, b, lib = self.config, self._b, _lib.lib()
        start = time.time()
        per_level = cfg.search_time_limit / (self.nesting_level + 1)
        budget = self._budget()
        timed = self.max_sims_per_level is None
        n_sims_total = 0

        # ---- level 0: grow the other agent's tree from the histories the root belief holds
        n0 = self._P1[:, self._SW]
        segs, off, sel_root, sel_off, sel_cnt = [], 0, [], [], []
        offsets = {}
        for u, bel in self._t0_beliefs.items():
            if bel.shape[0] > 0:
                offsets[u] = (off, bel.shape[0])
                segs.append(bel)
                off += bel.shape[0]
        for u in n0.cpu().numpy().tolist():
            if u in offsets:
                sel_root.append(u); sel_off.append(offsets[u][0]); sel_cnt.append(offsets[u][1])
        if sel_root:
            pcat = torch.cat(segs, dim=0).contiguous()
            d_root = torch.tensor(sel_root, dtype=torch.int32, device=b.device)
            d_off = torch.tensor(sel_off, dtype=torch.int64, device=b.device)
            d_cnt = torch.tensor(sel_cnt, dtype=torch.int32, device=b.device)
            def launch0(w, k, sim0):
                self._params0.n_waves = k
                _lib.check(lib.posg_search_multi(
                    b._handle, _lib.ptr(self._arena0), self._cap_nodes, self._cap_log, self._root_t,
                    _lib.ptr(pcat), pcat.shape[0], _lib.ptr(d_root), _lib.ptr(d_off), _lib.ptr(d_cnt),
                    len(sel_root), w, sim0, C.byref(self._params0), self._stream()))

            # the visits of a level-0 simulation spread over the histories of the root belief: the schedule is
            # driven by the simulations this tree has had so far
            done = self._run_waves(launch0, budget, per_level if timed else None, self._sim0, self._sim0)
            self._sim0 += done
            n_sims_total += done

        # ---- level 1: search the ego tree, the other agent acting by its tree.  This tree is
        # compacted at every re-root, so its budget is whatever room is left in the arena.
        if timed:
            budget = max(1024, self._cap_nodes - self._nodes1 - 4096)
        def launch1(w, k, sim0):
            self._params1.n_waves = k
            _lib.check(lib.posg_search_nested(
                self._nested, _lib.ptr(self._arena1), self._cap_nodes, self._cap_log, _lib.ptr(self._arena0),
                self._cap_nodes, self._cap_log, self._other, self._root1, self._root_t, _lib.ptr(self._P1),
                self._P1.shape[0], w, sim0, C.byref(self._params1), self._stream()))

        root_visits = int(self._stats(1, self._root1)["node_visits"])
        done = self._run_waves(launch1, budget, per_level if timed else None, self._sim1, root_visits)
        self._sim1 += done
        n_sims_total += done

        st = self._stats(1, self._root1)
        self._nodes1 = int(min(st["node_count"], self._cap_nodes))
        self._min_value, self._max_value = st["min_value"], st["max_value"]
        self.step_statistics.update({
            "search_time": time.time() - start, "search_depth": int(st["max_depth"]), "num_sims": n_sims_total,
            "min_value": self._min_value, "max_value": self._max_value})
        self._last_root_stats = st
        done_v, total = st["done"], st["total"]
        actions = self.action_spaces[self.agent_id]
        if cfg.action_selection == "pucb":
            if done_v.sum() == 0:
                return random.choice(actions)
            best = done_v.max()
            return random.choice([a for a in actions if done_v[a] == best])
        values = np.where(done_v > 0, total / np.maximum(done_v, 1), 0.0)
        best = values.max()
        return random.choice([a for a in actions if values[a] == best])

    def _run_waves(self, launch, budget: int, time_limit, sim0: int, visits: int) -> int:
        """Runs chained search waves (``launch(wave, n_waves, first simulation id)``) until ``budget`` simulations
        or ``time_limit`` seconds: a fixed ``wave_size`` when one was given, else the schedule that keeps the
        wave-parallel search equivalent to the sequential one (MCTS._scheduled_wave: one warp, two warps, then
        1/2048 of the visits the tree already has)."""
        from posggym_baselines_b200.planning.mcts import MCTS

        done, chunk, t0 = 0, 8, time.time()
        with self._dev():
            while done < budget:
                if self.wave_size:
                    wave = self.wave_size
                else:
                    floor = 32 if visits < 8192 else 64
                    wave = max(floor, min(self.max_wave, (int(visits * MCTS.WAVE_GROWTH) // 32) * 32))
                left = budget - done
                if left < wave:
                    wave, waves = left, 1
                else:
                    waves = max(1, min(chunk, left // wave))
                t_chunk = time.time()
                launch(wave, waves, sim0 + done)
                done += wave * waves
                visits += wave * waves
                torch.cuda.current_stream().synchronize()
                now = time.time()
                if time_limit is not None:
                    if now - t0 >= time_limit:
                        break
                    per_wave = (now - t_chunk) / waves
                    chunk = int(max(1, min(4096, 0.002 / max(per_wave, 1e-6), (time_limit - (now - t0)) / max(per_wave, 1e-6))))
                else:
                    chunk = 256
        return done

    def close(self):
        for p in self.search_policies.values():
            p.close()
        if self._nested is not None and self._nested.value:
            _lib.lib().posg_model_destroy(self._nested)
            self._nested = None
        self._arena1 = self._arena0 = self._arena1_spare = self._scratch = None
        self._P1 = None
        self._t0_beliefs = {}

    def __str__(self):
        return self.__class__.__name__
