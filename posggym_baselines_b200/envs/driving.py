"""Driving-v1 on the GPU: the batched model (``BatchedDriving``) and the scalar
``POSGModel``-shaped facade (``DrivingModel``) the reference planners consume
(call sites planning/mcts.py:181,191-192,333,418,
planning/belief.py:165).  All arithmetic runs in libposg_b200.so; this file only
marshals packed integer tensors (layouts: include/posg_b200.h).
"""
from __future__ import annotations

import ctypes as C
import random
from typing import Dict, List, NamedTuple, Optional, Sequence, Tuple

import numpy as np
import torch

from posggym_baselines_b200 import _lib
from posggym_baselines_b200.spaces import Discrete
from posggym_baselines_b200.spaces import Tuple as TupleSpace

# Layouts shipped by posggym's driving.py ('.' road, '#' block, '+' start, '-' destination,
# digits / a-j agent-specific start / destination).
GRIDS = {
    "14x14RoundAbout": (
        "#-..........-#\n"
        "-#####..#####+\n"
        ".#####..#####.\n"
        ".#####..#####.\n"
        ".###......###.\n"
        ".###......###.\n"
        "......##......\n"
        "......##......\n"
        ".###......###.\n"
        ".###......###.\n"
        ".#####..#####.\n"
        ".#####..#####.\n"
        "-#####..#####+\n"
        "#+..........+#\n"
    ),
    "7x7RoundAbout": (
        "#-...-#\n"
        "-##.##+\n"
        ".#...#.\n"
        "...#...\n"
        ".#...#.\n"
        "-##.##+\n"
        "#+...+#\n"
    ),
    "6x6Intersection": (
        "##0b##\n"
        "##..##\n"
        "d....3\n"
        "2....c\n"
        "##..##\n"
        "##a1##\n"
    ),
}

NUM_ACTIONS = 5
MAX_EPISODE_STEPS = 50
FLAG_ALL_DONE = 0x80


class VehicleState(NamedTuple):
    coord: Tuple[int, int]
    facing_dir: int
    speed: int
    dest_coord: Tuple[int, int]
    dest_reached: int
    crashed: int
    min_dest_dist: int
    init_dest_dist: int


class StepOut(NamedTuple):
    state: torch.Tensor     # int32 [n, A]
    obs: torch.Tensor       # int64 [n, A]
    rewards: torch.Tensor   # float64 [n, A]
    flags: torch.Tensor     # uint8 [n]
    actions: Optional[torch.Tensor]  # uint8 [n, A]


class BatchedDriving:
    """Batched Driving-v1 model; one packed int32 word per vehicle."""

    env_id = "Driving-v1"

    def __init__(
        self,
        grid: str = "14x14RoundAbout",
        num_agents: int = 2,
        obs_dim: Sequence[int] = (3, 1, 1),
        obstacle_collisions: bool = False,
        device: Optional[torch.device] = None,
    ):
        self.grid_name = grid
        self.grid_str = GRIDS[grid] if grid in GRIDS else grid
        self.num_agents = int(num_agents)
        self.obs_dim = tuple(int(v) for v in obs_dim)
        self.obstacle_collisions = bool(obstacle_collisions)
        rows = [r for r in self.grid_str.splitlines() if r]
        self.width, self.height = len(rows[0]), len(rows)
        self.dest_coords = [(x, y) for y, r in enumerate(rows) for x, ch in enumerate(r) if ch == "-" or ch in "abcdefghij"]
        # per-agent start / destination cells in parse order ('+' / '-': every agent; digit / letter: that agent)
        self.agent_starts = [[(x, y) for y, r in enumerate(rows) for x, ch in enumerate(r) if ch == "+" or ch == str(i)]
                             for i in range(self.num_agents)]
        self.agent_dests = [[(x, y) for y, r in enumerate(rows) for x, ch in enumerate(r) if ch == "-" or ch == "abcdefghij"[i]]
                            for i in range(self.num_agents)]
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
        if device is None:
            raise _lib.PosgB200Error("no CUDA device: posggym_baselines_b200 has no CPU fallback")
        self.device = torch.device(device)
        self._handle = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(
                _lib.lib().posg_driving_create(
                    self.grid_str.encode(), self.num_agents, *self.obs_dim,
                    int(self.obstacle_collisions), C.byref(self._handle),
                )
            )

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and h.value:
            try:
                _lib.lib().posg_model_destroy(h)
            except Exception:
                pass
            self._handle = None

    def initial_pick_bounds(self, picks_so_far: Sequence[int]) -> int:
        """Number of choices of the next initial-state pick given the earlier ones (per agent: a start cell among
        those still free, then a destination among those still free and different from the start) -- the bounds
        of the explicit ``picks`` of ``sample_initial_state``."""
        i, is_dest = divmod(len(picks_so_far), 2)
        used_s, used_d = set(), set()
        for j in range(i):
            s = [c for c in self.agent_starts[j] if c not in used_s][picks_so_far[2 * j]]
            used_s.add(s)
            used_d.add([c for c in self.agent_dests[j] if c not in used_d and c != s][picks_so_far[2 * j + 1]])
        avail = [c for c in self.agent_starts[i] if c not in used_s]
        if not is_dest:
            return len(avail)
        s = avail[picks_so_far[2 * i]]
        return len([c for c in self.agent_dests[i] if c not in used_d and c != s])

    num_picks = property(lambda self: 2 * self.num_agents)

    # ---- batched model calls -------------------------------------------------------
    def _empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def sample_initial_state(self, n: int, seed: int = 0, particle_offset: int = 0,
                             picks: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        state = out if out is not None else self._empty((n, self.num_agents), torch.int32)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().posg_driving_sample_initial_state(
                self._handle, n, _lib.ptr(picks), seed, particle_offset, _lib.ptr(state), _lib.stream_ptr()))
        return state

    def observe(self, state: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        n = state.shape[0]
        obs = out if out is not None else self._empty((n, self.num_agents), torch.int64)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().posg_driving_observe(self._handle, n, _lib.ptr(state), _lib.ptr(obs), _lib.stream_ptr()))
        return obs

    def step(self, state: torch.Tensor, actions: Optional[torch.Tensor] = None,
             order: Optional[torch.Tensor] = None, seed: int = 0, step_idx: int = 0,
             particle_offset: int = 0, out: Optional[StepOut] = None,
             live_count: Optional[torch.Tensor] = None, want_actions: bool = True) -> StepOut:
        n, a = state.shape[0], self.num_agents
        if out is None:
            out = StepOut(
                self._empty((n, a), torch.int32), self._empty((n, a), torch.int64),
                self._empty((n, a), torch.float64), self._empty((n,), torch.uint8),
                self._empty((n, a), torch.uint8) if (actions is None and want_actions) else None,
            )
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().posg_driving_step(
                self._handle, n, _lib.ptr(state), _lib.ptr(actions), _lib.ptr(order), seed, step_idx,
                particle_offset, _lib.ptr(out.state), _lib.ptr(out.obs), _lib.ptr(out.rewards),
                _lib.ptr(out.flags), _lib.ptr(out.actions), _lib.ptr(live_count), _lib.stream_ptr()))
        return out

    def step_host(self, state: np.ndarray, actions: np.ndarray, order: np.ndarray):
        """``model.step`` for a few rows in host memory (``posg_model_step_host``): numpy in, numpy out --
        ``(state [n, A] int32, obs [n, A] int64, rewards [n, A] float64, flags [n] uint8)``.  The scalar facade and
        the planners' plugin path step through this."""
        n, a = state.shape[0], self.num_agents
        state = np.ascontiguousarray(state, dtype=np.int32)
        actions = np.ascontiguousarray(actions, dtype=np.uint8)
        order = np.ascontiguousarray(order, dtype=np.uint8).reshape(n)
        out = (np.empty((n, a), np.int32), np.empty((n, a), np.int64), np.empty((n, a), np.float64),
               np.empty((n,), np.uint8))
        _lib.check(_lib.lib().posg_model_step_host(self._handle, n, _lib.ptr(state), _lib.ptr(actions), _lib.ptr(order),
                                                   1, *(_lib.ptr(o) for o in out)))
        return out

    def rollout(self, state: torch.Tensor, ego_agent: int, max_steps: int, discount: float,
                seed: int = 0, step0: int = 0, particle_offset: int = 0, want_state: bool = False,
                live_count: Optional[torch.Tensor] = None):
        n = state.shape[0]
        returns = self._empty((n,), torch.float64)
        steps = self._empty((n,), torch.uint8)
        final = self._empty((n, self.num_agents), torch.int32) if want_state else None
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().posg_driving_rollout(
                self._handle, n, _lib.ptr(state), ego_agent, max_steps, discount, seed, step0,
                particle_offset, _lib.ptr(returns), _lib.ptr(steps), _lib.ptr(final),
                _lib.ptr(live_count), _lib.stream_ptr()))
        return returns, steps, final

    # ---- host-buffer entry points (bench e2e) -----------------------------------------
    def rollout_host(self, state_host: np.ndarray, ego_agent: int, max_steps: int, discount: float,
                     seed: int = 0, step0: int = 0, particle_offset: int = 0,
                     returns_out: Optional[np.ndarray] = None, steps_out: Optional[np.ndarray] = None):
        n = state_host.shape[0]
        returns = returns_out if returns_out is not None else np.empty(n, np.float64)
        steps = steps_out if steps_out is not None else np.empty(n, np.uint8)
        live = np.zeros(1, np.uint64)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().posg_driving_rollout_host(
                self._handle, n, _lib.ptr(state_host), ego_agent, max_steps, discount, seed, step0,
                particle_offset, _lib.ptr(returns), _lib.ptr(steps), None, _lib.ptr(live)))
        return returns, steps, int(live[0])

    def sweep_host(self, state_host: np.ndarray, num_steps: int, seed: int = 0, particle_offset: int = 0,
                   outs: Optional[tuple] = None):
        n, a = state_host.shape[0], self.num_agents
        if outs is None:
            outs = (np.empty((n, a), np.int32), np.empty((n, a), np.int64),
                    np.empty((n, a), np.float64), np.empty(n, np.uint8))
        live = np.zeros(1, np.uint64)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().posg_driving_sweep_host(
                self._handle, n, _lib.ptr(state_host), num_steps, seed, particle_offset,
                _lib.ptr(outs[0]), _lib.ptr(outs[1]), _lib.ptr(outs[2]), _lib.ptr(outs[3]), _lib.ptr(live)))
        return outs, int(live[0])

    def sweep_host_compact(self, state_host: np.ndarray, num_steps: int, seed: int = 0, particle_offset: int = 0,
                           outs: Optional[tuple] = None):
        """``sweep_host`` with the rewards as one code byte per env (``decode_rewards`` expands them to the same
        float64 values): ``outs`` = (state [n, A] i32, obs [n, A] i64, reward code [n] u8, flags [n] u8)."""
        n, a = state_host.shape[0], self.num_agents
        if outs is None:
            outs = (np.empty((n, a), np.int32), np.empty((n, a), np.int64), np.empty(n, np.uint8), np.empty(n, np.uint8))
        live = np.zeros(1, np.uint64)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().posg_driving_sweep_host_compact(
                self._handle, n, _lib.ptr(state_host), num_steps, seed, particle_offset,
                _lib.ptr(outs[0]), _lib.ptr(outs[1]), _lib.ptr(outs[2]), _lib.ptr(outs[3]), _lib.ptr(live)))
        return outs, int(live[0])

    def decode_rewards(self, state_host: np.ndarray, code_host: np.ndarray, out: Optional[np.ndarray] = None):
        n = state_host.shape[0]
        out = out if out is not None else np.empty((n, self.num_agents), np.float64)
        _lib.check(_lib.lib().posg_driving_decode_rewards_host(self._handle, n, _lib.ptr(state_host), _lib.ptr(code_host),
                                                               _lib.ptr(out)))
        return out

    # ---- codec: packed words <-> posggym-style tuples -----------------------------------
    def pack_state(self, states: Sequence[Sequence[VehicleState]]) -> np.ndarray:
        out = np.zeros((len(states), self.num_agents), np.uint32)
        for i, st in enumerate(states):
            for a, v in enumerate(st):
                dest = self.dest_coords.index(tuple(v.dest_coord))
                if not 0 < int(v.init_dest_dist) < 64:
                    raise ValueError("init_dest_dist must be in 1..63 (6-bit field of the packed state)")
                out[i, a] = (
                    v.coord[0] | (v.coord[1] << 4) | (int(v.facing_dir) << 8) | (int(v.speed) << 10)
                    | (dest << 12) | (int(v.dest_reached) << 15) | (int(v.crashed) << 16)
                    | (int(v.min_dest_dist) << 17) | (int(v.init_dest_dist) << 24)
                )
        return out.view(np.int32)

    def unpack_state(self, packed) -> List[Tuple[VehicleState, ...]]:
        w = _to_numpy(packed).view(np.uint32).astype(np.int64)
        out = []
        for row in w:
            out.append(tuple(
                VehicleState(
                    (int(v & 15), int((v >> 4) & 15)), int((v >> 8) & 3), int((v >> 10) & 3),
                    self.dest_coords[int((v >> 12) & 7)], int((v >> 15) & 1), int((v >> 16) & 1),
                    int((v >> 17) & 127), int((v >> 24) & 63),
                )
                for v in row
            ))
        return out

    @staticmethod
    def pack_obs(obs) -> np.ndarray:
        """One posggym observation tuple -> packed key words (uint64[1])."""
        cells, speed, coord, dest, reached, crashed = obs
        o = 0
        for c, v in enumerate(cells):
            o |= int(v) << (2 * c)
        o |= int(speed) << 30
        o |= (int(coord[0]) | (int(coord[1]) << 4) | (int(dest[0]) << 8) | (int(dest[1]) << 12)
              | (int(reached) << 16) | (int(crashed) << 17)) << 32
        return np.array([o], dtype=np.uint64)

    @staticmethod
    def unpack_obs(packed) -> List[tuple]:
        """[n, A] packed words -> per env a tuple of per-agent posggym observations
        ``(cells[15], speed, (x, y), (dest x, dest y), dest_reached, crashed)``."""
        w = _to_numpy(packed).astype(np.int64)
        out = []
        for row in w:
            obs = []
            for o in row:
                o = int(o)
                cells = tuple((o >> (2 * c)) & 3 for c in range(15))
                obs.append((cells, (o >> 30) & 3, ((o >> 32) & 15, (o >> 36) & 15),
                            ((o >> 40) & 15, (o >> 44) & 15), (o >> 48) & 1, (o >> 49) & 1))
            out.append(tuple(obs))
        return out


def _to_numpy(t) -> np.ndarray:
    if isinstance(t, torch.Tensor):
        return t.detach().cpu().numpy()
    return np.asarray(t)


class _Spec:
    def __init__(self, id: str, max_episode_steps: Optional[int]):
        self.id = id
        self.max_episode_steps = max_episode_steps


class JointTimestep(NamedTuple):
    state: tuple
    observations: Dict[str, tuple]
    rewards: Dict[str, float]
    terminations: Dict[str, bool]
    truncations: Dict[str, bool]
    all_done: bool
    infos: Dict[str, dict]


class DrivingModel:
    """Scalar ``posggym.model.POSGModel`` facade over the batched kernels (drop-in for
    ``posggym.make("Driving-v1").model`` on the planning path).  One call = one kernel
    launch over a batch of one; planners that want throughput use ``.batched``."""

    is_symmetric = True

    def __init__(self, grid: str = "14x14RoundAbout", num_agents: int = 2,
                 obs_dim: Sequence[int] = (3, 1, 1), obstacle_collisions: bool = False,
                 device: Optional[torch.device] = None):
        self.batched = BatchedDriving(grid, num_agents, obs_dim, obstacle_collisions, device)
        self.possible_agents = tuple(str(i) for i in range(num_agents))
        self.action_spaces = {i: Discrete(NUM_ACTIONS) for i in self.possible_agents}
        w, h = self.batched.width, self.batched.height
        self.observation_spaces = {
            i: TupleSpace((
                TupleSpace(tuple(Discrete(4) for _ in range(15))), Discrete(4),
                TupleSpace((Discrete(w), Discrete(h))), TupleSpace((Discrete(w), Discrete(h))),
                Discrete(2), Discrete(2),
            ))
            for i in self.possible_agents
        }
        self.spec = _Spec("Driving-v1", MAX_EPISODE_STEPS)
        self._rng = random.Random()

    @property
    def rng(self) -> random.Random:
        return self._rng

    def seed(self, seed: Optional[int] = None):
        self._rng = random.Random(seed)
        for k, sp in enumerate(self.action_spaces.values()):
            sp.seed(None if seed is None else seed + 1 + k)

    def get_agents(self, state) -> List[str]:
        return list(self.possible_agents)

    def sample_initial_state(self):
        # the picks come from model.rng one by one (like posggym's model draws its start / destination cells);
        # the state is built by the kernel from the explicit picks
        b = self.batched
        picks: List[int] = []
        for _ in range(b.num_picks):
            picks.append(self._rng.randrange(b.initial_pick_bounds(picks)))
        st = b.sample_initial_state(1, picks=torch.tensor([picks], dtype=torch.uint8, device=b.device))
        return b.unpack_state(st)[0]

    def sample_initial_obs(self, state) -> Dict[str, tuple]:
        b = self.batched
        packed = torch.from_numpy(b.pack_state([state])).to(b.device)
        obs = b.unpack_obs(b.observe(packed))[0]
        return {i: obs[k] for k, i in enumerate(self.possible_agents)}

    def sample_agent_initial_state(self, agent_id: str, obs):
        raise NotImplementedError

    def host_step_noise(self, n: int) -> np.ndarray:
        """Explicit noise for ``n`` rows of ``batched.step_host`` / ``batched.step``: the processing-order draw of
        ``step``, n times (uint8 [n])."""
        import math

        k = math.factorial(len(self.possible_agents))
        return np.array([self._rng.randrange(k) for _ in range(n)], dtype=np.uint8)

    def batched_step_noise(self, n: int):
        return torch.from_numpy(self.host_step_noise(n)).to(self.batched.device)

    def step(self, state, actions: Dict[str, int]) -> JointTimestep:
        b = self.batched
        act = np.array([[int(actions[i]) for i in self.possible_agents]], dtype=np.uint8)
        codec = self.__dict__.get("_codec")
        if codec is None:
            from posggym_baselines_b200.envs.base import ScalarCodec

            codec = self._codec = ScalarCodec(b)
        st, ob, rw, fl = b.step_host(codec.pack(state), act, self.host_step_noise(1))
        nstate = codec.state(st)
        obs = codec.obs(ob[0])
        rew = rw[0]
        flags = int(fl[0])
        ids = self.possible_agents
        return JointTimestep(
            nstate,
            {i: obs[k] for k, i in enumerate(ids)},
            {i: float(rew[k]) for k, i in enumerate(ids)},
            {i: bool((flags >> k) & 1) for k, i in enumerate(ids)},
            {i: False for i in ids},
            bool(flags & FLAG_ALL_DONE),
            {i: {} for i in ids},
        )
