"""``MCTS`` / ``POMCP`` planner facade over the device search tree.

Same surface as the reference engine (/root/reference/posggym_baselines/planning/
mcts.py:22-739, pomcp.py:9-35): ``reset()``, ``step(obs) -> action``, ``update(action,
obs)``, ``get_action()``, ``close()``, ``step_statistics`` (keys of mcts.py:140-153),
``stat_tracker``, ``root.belief.size()``.  The tree, the particle beliefs, the descents,
leaf rollouts, backups, the initial-belief rejection loop and reinvigoration all run in
libposg_b200.so; this file only sequences launches and applies the wall-clock budget.

Device search = the POMCP configuration: ``RandomSearchPolicy`` for the ego agent and
``RandomOtherAgentPolicy`` for the others (either value of ``state_belief_only``; ``truncated`` is
irrelevant because ``RandomSearchPolicy.get_value`` raises: rollout fallback, mcts.py:392-399).  Any
other ``SearchPolicy`` / ``OtherAgentPolicy`` plugin runs through its scalar methods on the host
(planning/host_search.py) with the model arithmetic on the device.

Extra, B200-only knobs are keyword arguments of the constructor, never config fields:
``max_sims`` (stop after that many simulations instead of / as well as the time limit),
``wave_size`` (simulations per launch; 1 = the reference's strictly sequential search),
``root_parallel`` (merge root statistics over torch.distributed ranks: one allreduce per
decision), ``trees_per_gpu`` (root parallelisation inside the GPU: that many independent trees, one CTA / SM
each, searched by ONE launch per chunk and merged at the root like the cross-GPU allreduce; every tree is the
single-tree search), ``node_capacity`` / ``log_capacity`` (size of the device arenas, shared by the trees).  Like the
reference (mcts.py:261-263) the planner keeps the whole subtree under the new root -- its
statistics, child lists and the particles logged at its nodes -- and drops the rest: at
every re-root the subtree is copied, renumbered, into a second arena (posg_tree_reroot).
"""
from __future__ import annotations

import ctypes as C
import logging
import math
import random
import time
from typing import Dict, Optional

import numpy as np
import torch

from posggym_baselines_b200 import _lib
from posggym_baselines_b200.planning.config import MCTSConfig
from posggym_baselines_b200.planning.other_policy import OtherAgentPolicy, RandomOtherAgentPolicy
from posggym_baselines_b200.planning.search_policy import RandomSearchPolicy, SearchPolicy
from posggym_baselines_b200.planning.utils import PlanningStatTracker

_SELECTION = {"ucb": 0, "pucb": 1, "uniform": 2}


class _BeliefView:
    """``planner.root.belief`` stand-in: ``size()`` and ``particles`` (packed states)."""

    def __init__(self, planner):
        self._p = planner

    def size(self) -> int:
        t = self._p._root_particles
        return 0 if t is None else int(t.shape[0])

    @property
    def particles(self):
        t = self._p._root_particles
        return [] if t is None else self._p.model.batched.unpack_state(t)


class _RootView:
    def __init__(self, planner):
        self._p = planner
        self.belief = _BeliefView(planner)

    @property
    def t(self) -> int:
        return self._p._root_t

    @property
    def is_absorbing(self) -> bool:
        return self._p._root_absorbing

    @property
    def visits(self) -> int:
        return int(self._p._stats()["node_visits"]) if self._p._arena is not None else 0


class MCTS:
    """Partially observable multi-agent Monte-Carlo planning on the GPU."""

    #: largest wave as a fraction of the root's visits for which the wave-parallel search reproduces the
    #: reference planner's root statistics at large budgets (tests/test_pomcp_gpu.py)
    WAVE_GROWTH = 1.0 / 2048.0
    #: smallest wave by root visits: (visits from which it applies, wave size), ascending
    WAVE_FLOORS = ((0, 32), (8192, 64))
    #: host threads that run the per-tree updates of a forest (trees_per_gpu >= 4)
    UPDATE_THREADS = 8

    def __init__(
        self,
        model,
        agent_id: str,
        config: MCTSConfig,
        other_agent_policies: Dict[str, OtherAgentPolicy],
        search_policy: SearchPolicy,
        *,
        max_sims: Optional[int] = None,
        wave_size: Optional[int] = None,
        max_wave: int = 262144,
        wave_growth: float = 0.0,
        node_capacity: int = 1 << 27,
        log_capacity: int = 1 << 28,
        root_parallel: bool = False,
        trees_per_gpu: int = 1,
        _forest_index: int = 0,
    ):
        self.model = model
        self.agent_id = agent_id
        self.config = config
        self.num_agents = len(model.possible_agents)
        assert len(other_agent_policies) == self.num_agents - 1
        assert agent_id not in other_agent_policies
        self.other_agent_policies = other_agent_policies
        self.search_policy = search_policy
        # RandomSearchPolicy / RandomOtherAgentPolicy are executed inside the kernels; any other plugin is a
        # stateful Python object, so the descent that calls its scalar methods runs on the host
        # (planning/host_search.py) with the model arithmetic still on the device
        self._host_mode = not (type(search_policy) is RandomSearchPolicy and all(
            type(p) is RandomOtherAgentPolicy for p in other_agent_policies.values()))
        # state_belief_only=False (mcts.py:344-354) makes particles carry the joint history and the other agents'
        # policy states.  With RandomOtherAgentPolicy the policy state is empty and nothing the search computes
        # reads the history, so the device search is the same; the histories are not materialised
        # (planner.root.belief.particles shows states).
        if not hasattr(model, "batched"):
            raise TypeError("model must be a posggym_baselines_b200 model facade (needs `.batched`)")
        self._b = model.batched
        self._ego = list(model.possible_agents).index(agent_id)
        self.action_space = list(range(model.action_spaces[agent_id].n))

        if config.step_limit is not None:
            self.step_limit = config.step_limit
        elif model.spec is not None:
            self.step_limit = model.spec.max_episode_steps
        else:
            self.step_limit = float("inf")

        self._rng = random.Random(config.seed)
        self._seed = config.seed if config.seed is not None else random.getrandbits(48)
        self.max_sims = max_sims
        self.wave_size = wave_size
        self.max_wave = max_wave
        self.wave_growth = float(wave_growth) if wave_growth else self.WAVE_GROWTH
        self.root_parallel = root_parallel
        self.allreduce_time = 0.0
        self._seed_shared = config.seed if config.seed is not None else 0
        self._cap_nodes, self._cap_log = int(node_capacity), int(log_capacity)
        self._forest_index = int(_forest_index)
        # tree k of a forest draws its belief particles from its own Philox sub-sequence
        self._draw_base = self._forest_index << 48
        trees_per_gpu = int(trees_per_gpu)
        if trees_per_gpu > 1:
            # the trees share the arena budget
            self._cap_nodes = max(1 << 20, self._cap_nodes // trees_per_gpu)
            self._cap_log = max(1 << 22, self._cap_log // trees_per_gpu)

        lib = _lib.lib()
        na, ow, sw = C.c_int(), C.c_int(), C.c_int()
        _lib.check(lib.posg_model_info(self._b._handle, C.byref(na), C.byref(ow), C.byref(sw)))
        self._A, self._OW, self._SW = na.value, ow.value, sw.value
        assert self._A == len(self.action_space)
        self._arena: Optional[torch.Tensor] = None
        self._bbuf, self._bflip, self._rej = [None, None], 0, None
        self._params = _lib.SearchParams()
        kb = config.known_bounds
        self._params.discount = config.discount
        self._params.c = config.c
        self._params.known_min = kb.min if kb else 1.0
        self._params.known_max = kb.max if kb else 0.0
        self._params.pucb_exploration_fraction = config.pucb_exploration_fraction
        self._params.depth_limit = config.depth_limit
        self._params.step_limit = int(min(self.step_limit, 250))
        self._params.selection = _SELECTION[config.action_selection]
        self._params.ego_agent = self._ego
        self._params.seed = self._seed & 0xFFFFFFFFFFFFFFFF
        if root_parallel:
            from posggym_baselines_b200 import parallel

            # same root belief on every rank (belief draws use the shared seed); the simulations
            # of rank r use the Philox sub-sequence starting at r * 2**40
            self._sim_base = parallel.world()[0] << 40
        else:
            self._sim_base = 0

        self._device_root = _RootView(self)
        self._logger = logging.getLogger()
        self.step_statistics: Dict[str, float] = {}
        self.stat_tracker = PlanningStatTracker(self)
        self._siblings = []
        self._update_pool = None
        if trees_per_gpu > 1:
            if self._host_mode:
                raise NotImplementedError("trees_per_gpu > 1 needs the device search (random search / other-agent policies)")
            self._siblings = [
                MCTS(model, agent_id, config, other_agent_policies, search_policy, max_sims=max_sims,
                     wave_size=wave_size, max_wave=max_wave, wave_growth=wave_growth, node_capacity=self._cap_nodes,
                     log_capacity=self._cap_log, root_parallel=False, _forest_index=k)
                for k in range(1, trees_per_gpu)]
            for sib in self._siblings:
                sib._seed, sib._params.seed = self._seed, self._params.seed   # one Philox key, disjoint counters
            with torch.cuda.device(self._b.device):
                _lib.check(lib.posg_search_warmup(self._b._handle))   # no lazy kernel loads in the threaded update
        self._host = None
        if self._host_mode:
            self._reset_step_statistics_host()
            self._host = self._host_engine()(self)
        self.reset()

    def _host_engine(self):
        """Engine class of the plugin path (typed planners substitute their own)."""
        from posggym_baselines_b200.planning.host_search import HostSearch

        return HostSearch

    @property
    def root(self):
        """``planner.root``: ``belief.size()`` / ``belief.particles``, ``t``, ``is_absorbing``, ``visits``."""
        return self._host.root if self._host is not None else self._device_root

    def _reset_step_statistics_host(self):
        self._min_value, self._max_value = float("inf"), -float("inf")
        self._reset_step_statistics()

    def _device_initial_belief(self, init_obs, need: int) -> torch.Tensor:
        """MCTS._initial_update's rejection loop (mcts.py:186-194) in one launch: ``need`` packed states whose
        first ego observation is ``init_obs``."""
        b = self._b
        out = torch.empty((need, self._SW), dtype=torch.int32, device=b.device)
        res = np.zeros(3, np.int64)
        key = np.ascontiguousarray(b.pack_obs(init_obs), dtype=np.uint64)
        max_attempts = 1 << 26
        with torch.cuda.device(b.device):
            _lib.check(_lib.lib().posg_belief_initial(
                b._handle, self._ego, _lib.ptr(key), need, max_attempts, self._params.seed,
                self._next_draw(max_attempts), _lib.ptr(out), _lib.ptr(res), self._stream()))
        if res[0] < need:
            raise RuntimeError(f"initial belief: only {res[0]} of {need} particles matched the first observation")
        return out

    # ------------------------------------------------------------------ plumbing
    def _stream(self):
        return _lib.stream_ptr()

    def _ensure_arena(self):
        if self._arena is None:
            lib = _lib.lib()
            nbytes = lib.posg_tree_arena_bytes(self._b._handle, self._cap_nodes, self._cap_log)
            _lib.check(int(nbytes))
            self._arena = torch.empty(int(nbytes), dtype=torch.uint8, device=self._b.device)
            self._arena_spare = torch.empty(int(nbytes), dtype=torch.uint8, device=self._b.device)
            # the spare must be clean once: posg_tree_reroot only re-initialises the slots its previous life used
            with torch.cuda.device(self._b.device):
                _lib.check(lib.posg_tree_reset(self._b._handle, _lib.ptr(self._arena_spare), self._cap_nodes,
                                               self._cap_log, 0, 1.0, 0.0, self._stream()))
            self._reroot_scratch = torch.empty(3 * self._cap_nodes, dtype=torch.int32, device=self._b.device)
            # belief buffers of the update, sized for the whole particle log up front: an allocation during the
            # update (cudaMalloc behind the caching allocator, ~0.1 s with tens of GB of arenas mapped) was the
            # source of every update-time spike measured
            rows = min(self._cap_log, 1 << 26) + self.config.num_particles + self.config.extra_particles
            cap = 1 << max(12, int(rows - 1).bit_length())
            self._bbuf = [torch.empty((cap, self._SW), dtype=torch.int32, device=self._b.device) for _ in range(2)]
            self._rej = torch.empty((max(1, self.config.num_particles + self.config.extra_particles), self._SW),
                                    dtype=torch.int32, device=self._b.device)

    def _stats(self, node: Optional[int] = None) -> Dict[str, object]:
        out = np.zeros(3 * self._A + 10, np.float64)
        with torch.cuda.device(self._b.device):
            _lib.check(_lib.lib().posg_tree_node_stats(
                self._b._handle, _lib.ptr(self._arena), self._cap_nodes, self._cap_log,
                self._root if node is None else node, _lib.ptr(out), self._stream()))
        return self._unpack_stats(out)

    # ------------------------------------------------------------------ step / reset
    def step(self, obs):
        assert self.root.t <= self.step_limit
        if self.root.is_absorbing:
            for k in self.step_statistics:
                self.step_statistics[k] = np.nan
            return self._last_action
        self._reset_step_statistics()
        self.update(self._last_action, obs)
        self._last_action = self.get_action()
        self._step_num += 1
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

    def reset(self):
        self.stat_tracker.reset_episode()
        if self._host is not None:
            self._host.reset()
        for sib in self._siblings:
            sib.reset()
        self._step_num = 0
        self._root = 0
        self._root_t = 0
        self._root_absorbing = False
        self._root_particles: Optional[torch.Tensor] = None
        self._last_action = None
        self._sim_counter = 0
        self._draw_counter = 0
        self._nodes_bound = 1
        self._log_bound = 0
        self._min_value, self._max_value = float("inf"), -float("inf")
        self.tree_compactions = 0
        self.search_info = {}
        self._reset_step_statistics()

    def _reset_step_statistics(self):
        self.step_statistics = {
            "search_time": 0.0, "update_time": 0.0, "reinvigoration_time": 0.0, "evaluation_time": 0.0,
            "policy_calls": 0, "inference_time": 0.0, "search_depth": 0, "num_sims": 0, "mem_usage": 0,
            "min_value": self._min_value, "max_value": self._max_value,
        }

    # ------------------------------------------------------------------ update
    def update(self, action, obs):
        if self._host is not None:
            start = time.time()
            if self._host.root.t == 0:
                self._last_action = None
            self._host.update(action, obs)
            self.step_statistics["update_time"] = time.time() - start
            return
        if self._root_absorbing:
            return
        start = time.time()

        def update_tree(tree):
            if tree._root_absorbing:
                return
            if tree._root_t == 0:
                tree._last_action = None
                tree._initial_update(obs)
            else:
                tree._update(action, obs)

        trees = [self] + self._siblings
        if len(trees) < 4:
            for tree in trees:
                update_tree(tree)
        else:
            # the per-tree update is a handful of tiny launches with a host synchronisation after each: run the
            # trees on a few host threads, each with its own stream (the C-ABI calls release the GIL while they wait)
            if self._update_pool is None:
                from concurrent.futures import ThreadPoolExecutor

                n = min(self.UPDATE_THREADS, len(trees))
                self._update_pool = ThreadPoolExecutor(max_workers=n)
                self._update_streams = [torch.cuda.Stream(device=self._b.device) for _ in range(n)]
            torch.cuda.current_stream(self._b.device).synchronize()
            self._update_trace = []
            n = len(self._update_streams)

            def work(w):
                with torch.cuda.device(self._b.device), torch.cuda.stream(self._update_streams[w]):
                    for tree in trees[w::n]:
                        t_tree = time.time()
                        update_tree(tree)
                        self._update_trace.append((tree._forest_index, w, t_tree - start, time.time() - t_tree, getattr(tree, '_update_steps_us', None)))
                    torch.cuda.current_stream().synchronize()

            for f in [self._update_pool.submit(work, w) for w in range(n)]:
                f.result()
        torch.cuda.synchronize(self._b.device)
        for sib in self._siblings:
            self.step_statistics["reinvigoration_time"] += sib.step_statistics["reinvigoration_time"]
            sib.step_statistics["reinvigoration_time"] = 0.0
        self.step_statistics["update_time"] = time.time() - start

    def _next_draw(self, n: int) -> int:
        d = self._draw_counter
        self._draw_counter += int(n)
        return self._draw_base + d

    def _initial_update(self, init_obs):
        cfg, b = self.config, self._b
        self._ensure_arena()
        need = cfg.num_particles + cfg.extra_particles
        kb = cfg.known_bounds
        with torch.cuda.device(b.device):
            _lib.check(_lib.lib().posg_tree_reset(
                b._handle, _lib.ptr(self._arena), self._cap_nodes, self._cap_log, 1,
                kb.min if kb else 1.0, kb.max if kb else 0.0, self._stream()))
        out = self._device_initial_belief(init_obs, need)
        self._root, self._root_t, self._root_absorbing = 0, 1, False
        self._root_particles = out
        self._nodes_bound, self._log_bound = 1, 0

    def _update(self, action, obs):
        """``MCTS._update`` (mcts.py:229-263) in one library call (posg_tree_update): advance to the child, gather
        the particles the searches left there (belief.add_particle, mcts.py:371), reinvigorate, re-root keeping the
        subtree.  One call = one interpreter-lock hand-over per tree when a forest updates on several threads."""
        cfg, b = self.config, self._b
        key = np.ascontiguousarray(b.pack_obs(obs), dtype=np.uint64)
        target = cfg.num_particles + cfg.extra_particles
        # (the log's length is known on the host: get_action / the last re-root recorded it).  Two persistent belief
        # buffers, grown geometrically and used in turn (the old belief is the parent of the reinvigoration): no
        # allocator traffic in the update
        log_bound = min(self._log_bound, self._cap_log)
        rows = log_bound + target
        self._bflip ^= 1
        buf = self._bbuf[self._bflip]
        if buf is None or buf.shape[0] < rows:
            cap = 1 << max(12, int(rows - 1).bit_length())
            buf = self._bbuf[self._bflip] = torch.empty((cap, self._SW), dtype=torch.int32, device=b.device)
        if self._rej is None:
            self._rej = torch.empty((max(target, 1), self._SW), dtype=torch.int32, device=b.device)
        res = np.zeros(8, np.int64)
        t0 = time.time()
        with torch.cuda.device(b.device):
            _lib.check(_lib.lib().posg_tree_update(
                b._handle, _lib.ptr(self._arena), _lib.ptr(self._arena_spare), self._cap_nodes, self._cap_log,
                self._root, int(action), _lib.ptr(key), self._ego, _lib.ptr(self._root_particles),
                self._root_particles.shape[0], _lib.ptr(buf), buf.shape[0], log_bound, target,
                float(cfg.reinvigoration_sample_limit_factor), self._params.seed ^ 0x5DEECE66D,
                self._draw_base + self._draw_counter, _lib.ptr(self._rej), _lib.ptr(self._reroot_scratch),
                _lib.ptr(res), self._stream()))
        child, flags = int(res[0]), int(res[1])
        self._update_steps_us = [16 * ((int(res[7]) >> sh) & 0xFFFF) for sh in (0, 16, 32, 48)]   # advance, gather, reinvigorate, re-root
        if child < 0:
            raise RuntimeError("search tree arena is full; raise node_capacity")
        self._draw_counter += int(res[5])
        absorbing = bool(flags & 2)
        if not absorbing:
            self._root_particles = buf[:int(res[2])]
            if res[6]:
                self.step_statistics["reinvigoration_time"] += time.time() - t0   # (the whole call: it ran inside)
        self._root_t += 1
        self._root_absorbing = absorbing
        # the subtree under the new root was kept, renumbered into the other arena (mcts.py:261-263)
        self._arena, self._arena_spare = self._arena_spare, self._arena
        self._root = 0
        self._nodes_bound, self._log_bound = int(res[3]), int(res[4])
        self.tree_compactions += 1

    # ------------------------------------------------------------------ search
    def get_action(self):
        if self._host is not None:
            return self._host.get_action()
        if self._root_absorbing:
            return self.action_space[0]
        cfg, b = self.config, self._b
        start = time.time()
        n_sims = n_waves = wave_now = 0
        trees = [self] + [t for t in self._siblings if not t._root_absorbing]
        forest = len(trees) > 1
        A = self._A
        lib = _lib.lib()
        entries = None
        if forest:
            tab = np.zeros((len(trees), 4), np.int64)
            for k, t in enumerate(trees):
                tab[k] = (t._arena.data_ptr(), t._root_particles.data_ptr(), t._root_particles.shape[0], t._root)
            entries = torch.from_numpy(tab).to(b.device)
        nodes_before = sum(t._nodes_bound for t in trees)
        # Simulations of one wave see each other's virtual visits but not each other's returns, so a wave may
        # only be a small fraction of the visits the root already has: wave = wave_growth * N, at least 32 (one
        # warp).  At equal budgets the search then matches the reference's sequential loop (tests/
        # test_pomcp_gpu.py: 65 536 simulations against the unmodified reference planner; waves of N / 512 no
        # longer do).  The root keeps its subtree's visits across re-roots, so later decisions of an episode
        # start with larger waves.  Consecutive waves are chained inside one launch (posg_search_waves; with
        # trees_per_gpu > 1 posg_search_forest runs the same chain on every tree at once); the clock is read
        # between chunks of waves sized for ~2 ms.
        tree_visits = tree_sims = None
        if self.wave_size:
            root_visits = 0
        elif forest:
            # every tree follows the schedule with its OWN root visits (after a re-root the trees kept different
            # numbers of visits); the launch is sized for the largest wave
            tree_visits = self._forest_stats(entries, len(trees))[:, 3 * A].astype(np.int64)
            tree_sims = np.zeros(len(trees), np.int64)
            root_visits = int(tree_visits.min())
        else:
            root_visits = int(self._stats()["node_visits"])
        per_tree = forest and tree_visits is not None and self.max_sims is None
        tree_waves = None
        if per_tree:
            tree_room = np.array([max(32, ((self._cap_nodes - t._nodes_bound) // 2048) // 32 * 32) for t in trees], np.int64)
        chunk = 8          # waves per launch, adapted to ~2 ms per chunk
        with torch.cuda.device(b.device):
            stream = self._stream()
            while True:
                if self.wave_size:
                    wave = self.wave_size
                elif per_tree:
                    # (every simulation reserves a node id: a tree's wave is also held to what its arena can take in
                    # a decision of ~2 000 waves, so that no tree runs out of ids before the budget ends)
                    new_waves = np.minimum(np.minimum(self._scheduled_waves(tree_visits), tree_room), 1024)
                    if tree_waves is None or not np.array_equal(new_waves, tree_waves):   # (rarely changes)
                        tree_waves = new_waves
                        tab[:, 3] = tab[:, 3] & 0xFFFFFFFF | (tree_waves << 32)
                        entries.copy_(torch.from_numpy(tab), non_blocking=False)
                    wave = int(tree_waves.max())   # one CTA per tree
                else:
                    wave = self._scheduled_wave(root_visits)
                    if forest:
                        wave = min(wave, 1024)   # one CTA per tree
                waves = chunk
                if not self.wave_size and not per_tree:   # stay on this wave size only while the schedule allows it
                    grow_at = int((wave + 32) / self.wave_growth) + 1 if wave < self.max_wave else 1 << 62
                    for visits, _ in self.WAVE_FLOORS:   # the next floor of the schedule, if it comes first
                        if root_visits < visits < grow_at:
                            grow_at = visits
                    waves = max(1, min(chunk, (grow_at - root_visits + wave - 1) // wave))
                if self.max_sims is not None:
                    left = self.max_sims - n_sims
                    if left <= 0:
                        break
                    if left < wave:
                        wave, waves = left, 1
                    else:
                        waves = min(waves, left // wave)
                wave_now = wave
                t_chunk = time.time()
                # every simulation reserves one node id (none once the arena is full: the kernel then evaluates
                # unexpanded histories by rollouts without storing them)
                if forest:
                    _lib.check(lib.posg_search_forest(
                        b._handle, _lib.ptr(entries), len(trees), self._cap_nodes, self._cap_log, self._root_t, wave,
                        waves, self._sim_base + self._sim_counter, C.byref(self._params), stream))
                else:
                    self._nodes_bound = min(self._cap_nodes, self._nodes_bound + wave * waves)
                    _lib.check(lib.posg_search_waves(
                        b._handle, _lib.ptr(self._arena), self._cap_nodes, self._cap_log, self._root, self._root_t,
                        _lib.ptr(self._root_particles), self._root_particles.shape[0], wave, waves,
                        self._sim_base + self._sim_counter, C.byref(self._params), stream))
                self._sim_counter += wave * waves
                n_sims += wave * waves
                n_waves += waves
                root_visits += wave * waves
                if per_tree:
                    tree_sims += tree_waves * waves
                    tree_visits += tree_waves * waves
                torch.cuda.current_stream().synchronize()
                now = time.time()
                if self.max_sims is None and now - start >= cfg.search_time_limit:
                    break
                if self.max_sims is not None and cfg.search_time_limit < float("inf") and self._respect_time \
                        and now - start >= cfg.search_time_limit:
                    break
                dt = now - t_chunk
                if waves == chunk:
                    chunk = int(max(1, min(4096, chunk * 0.002 / max(dt, 1e-5))))
                if self.max_sims is None:   # never start a chunk that would overshoot the budget by more than ~2 ms
                    remaining = cfg.search_time_limit - (now - start)
                    per_wave = dt / max(1, waves)
                    chunk = int(max(1, min(chunk, remaining / max(per_wave, 1e-6))))
        if forest:
            rows = self._forest_stats(entries, len(trees))
            sts = [self._unpack_stats(r) for r in rows]
        else:
            sts = [self._stats()]
        for t, st_k in zip(trees, sts):
            t._nodes_bound, t._log_bound = int(st_k["node_count"]), int(st_k["log_count"])
            t._min_value, t._max_value = st_k["min_value"], st_k["max_value"]
            t._last_root_stats = st_k
        st = sts[0]
        n_total = int(tree_sims.sum()) if per_tree else n_sims * len(trees)
        if per_tree:
            n_sims = int(tree_sims.mean())
        self._min_value = min(s_["min_value"] for s_ in sts)
        self._max_value = max(s_["max_value"] for s_ in sts)
        self.step_statistics["search_time"] = time.time() - start
        self.step_statistics["search_depth"] = int(max(s_["max_depth"] for s_ in sts))
        self.step_statistics["num_sims"] = n_total
        self.step_statistics["min_value"] = self._min_value
        self.step_statistics["max_value"] = self._max_value
        # B200-only diagnostics (not a reference key): how much of the search could still grow the tree
        overflow = int(sum(s_["overflow"] for s_ in sts))
        node_count = int(sum(s_["node_count"] for s_ in sts))
        self.search_info = {
            "num_sims": n_total, "expanding_sims": n_total - overflow, "overflow_sims": overflow,
            "overflow_share": overflow / max(1, n_total), "nodes_created": int(sum(s_["nodes_created"] for s_ in sts)),
            "ids_reserved": node_count - nodes_before,
            "nodes_in_tree": node_count,
            "arena_fill": max(int(s_["node_count"]) for s_ in sts) / self._cap_nodes,
            "log_fill": min(1.0, max(int(s_["log_count"]) for s_ in sts) / self._cap_log), "waves": n_waves,
            "last_wave": wave_now, "trees": len(trees), "sims_per_tree": n_sims}
        done = sum(s_["done"] for s_ in sts)
        total = sum(s_["total"] for s_ in sts)
        return self._final_action(done, total, local_trees=len(trees))

    def _unpack_stats(self, out: np.ndarray) -> Dict[str, object]:
        A = self._A
        e = out[3 * A:]
        return {
            "visits": out[:A].copy(), "done": out[A:2 * A].copy(), "total": out[2 * A:3 * A].copy(),
            "node_visits": e[0], "node_count": e[1], "log_count": e[2], "max_depth": e[3],
            "min_value": e[4], "max_value": e[5], "overflow": e[6], "sims_done": e[7], "flags": int(e[8]),
            "nodes_created": e[9],
        }

    def _forest_stats(self, entries: torch.Tensor, n_trees: int) -> np.ndarray:
        out = np.zeros((n_trees, 3 * self._A + 10), np.float64)
        with torch.cuda.device(self._b.device):
            _lib.check(_lib.lib().posg_forest_stats(self._b._handle, _lib.ptr(entries), n_trees, self._cap_nodes,
                                                    self._cap_log, _lib.ptr(out), self._stream()))
        return out

    def _scheduled_wave(self, root_visits: int) -> int:
        """Simulations per wave for a root with ``root_visits`` visits: one warp while the tree is young, two
        warps from 8 192 visits on, then ``wave_growth`` (1/2048) of the visits -- the sizes for which the
        search reproduces the reference planner's root statistics at 65 536 and 262 144 simulations."""
        floor = 32
        for visits, wave in self.WAVE_FLOORS:
            if root_visits >= visits:
                floor = wave
        return max(floor, min(self.max_wave, (int(root_visits * self.wave_growth) // 32) * 32))

    def _scheduled_waves(self, root_visits: np.ndarray) -> np.ndarray:
        """``_scheduled_wave`` for an array of root visit counts (the trees of a forest)."""
        floors = np.full(root_visits.shape, 32, np.int64)
        for visits, wave in self.WAVE_FLOORS:
            floors = np.where(root_visits >= visits, wave, floors)
        grown = (root_visits * self.wave_growth).astype(np.int64) // 32 * 32
        return np.maximum(floors, np.minimum(self.max_wave, grown))

    _respect_time = False  # with max_sims set, the time limit only applies when asked for

    def _final_action(self, done, total, local_trees: int = 1):
        from posggym_baselines_b200 import parallel

        done, total = np.array(done, np.float64), np.array(total, np.float64)
        merged = False
        if self.root_parallel and parallel.world()[1] > 1:
            # root-parallel merge: one allreduce of |A| x {visits, total_value} per decision
            t0 = time.time()
            if local_trees > 1:
                buf = torch.from_numpy(np.concatenate([done, total])).to(self._b.device)
            else:
                buf = torch.empty(2 * self._A, dtype=torch.float64, device=self._b.device)
                with torch.cuda.device(self._b.device):
                    _lib.check(_lib.lib().posg_tree_root_stats_device(
                        self._b._handle, _lib.ptr(self._arena), self._cap_nodes, self._cap_log, self._root,
                        _lib.ptr(buf), self._stream()))
            done, total = parallel.merge_root_statistics_device(buf)
            self.allreduce_time = time.time() - t0
            merged = True
        self._merged_root = (done, total)
        if merged:
            return parallel.select_root_action(done, total, self.action_space, self.config.action_selection,
                                               (self._seed_shared << 8) ^ self._step_num)
        # single tree: the reference's selection with its global-`random` tie-break
        if self.config.action_selection == "pucb":
            # max_visit_action_selection (mcts.py:566-581)
            if done.sum() == 0:
                return random.choice(self.action_space)
            best = done.max()
            return random.choice([a for a in self.action_space if done[a] == best])
        # max_value_action_selection (mcts.py:583-600); unvisited action nodes hold value 0.0
        values = np.where(done > 0, total / np.maximum(done, 1), 0.0)
        best = values.max()
        return random.choice([a for a in self.action_space if values[a] == best])

    def root_action_values(self) -> np.ndarray:
        done, total = self._merged_root
        return np.where(done > 0, total / np.maximum(done, 1), 0.0)

    # ------------------------------------------------------------------ misc
    def close(self):
        self.search_policy.close()
        for policy in self.other_agent_policies.values():
            policy.close()
        self._arena = self._arena_spare = self._reroot_scratch = None
        self._root_particles = None
        for sib in self._siblings:
            sib._arena = sib._arena_spare = sib._reroot_scratch = None
            sib._root_particles = None
        self._siblings = []
        if self._update_pool is not None:
            self._update_pool.shutdown(wait=True)
            self._update_pool = None

    def __str__(self):
        return f"{self.__class__.__name__}"


class POMCP(MCTS):
    """POMCP = MCTS with random other agents and a belief over environment state only
    (reference: planning/pomcp.py:9-35)."""

    def __init__(self, model, agent_id: str, config: MCTSConfig, search_policy: SearchPolicy, **b200_options):
        other_agent_policies = {
            i: RandomOtherAgentPolicy(model, i) for i in model.possible_agents if i != agent_id
        }
        if not config.state_belief_only:
            # creates copy of config with state_belief_only=True
            config = config.replace(state_belief_only=True)
        super().__init__(model, agent_id, config, other_agent_policies, search_policy, **b200_options)
