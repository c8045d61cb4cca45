"""Model-generic batched wrapper over the ``posg_model_*`` C-ABI entry points and the
scalar ``POSGModel`` facade built on it."""
from __future__ import annotations

import ctypes as C
import random
from typing import Dict, List, NamedTuple, Optional

import numpy as np
import torch

from posggym_baselines_b200 import _lib

FLAG_ALL_DONE = 0x80


class StepOut(NamedTuple):
    state: torch.Tensor     # int32 [n, SW]
    obs: torch.Tensor       # int64 [n, A, OW]
    rewards: torch.Tensor   # float64 [n, A]
    flags: torch.Tensor     # uint8 [n]
    actions: Optional[torch.Tensor]  # uint8 [n, A]


class Spec:
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


class ScalarCodec:
    """Memo of the packed words <-> posggym value conversions of the SCALAR facade calls.  A planner's host path steps
    one state at a time and keeps seeing the same states and observations; the Python codec loops cost more than the
    device round trip, so their results are remembered by packed bytes / by (hashable) state."""

    LIMIT = 1 << 17

    def __init__(self, batched):
        self.b = batched
        self._packed: dict = {}
        self._state: dict = {}
        self._obs: dict = {}

    def pack(self, state) -> np.ndarray:
        row = self._packed.get(state)
        if row is None:
            if len(self._packed) >= self.LIMIT:
                self._packed.clear()
            row = self._packed[state] = self.b.pack_state([state])
        return row

    def state(self, packed_row: np.ndarray):
        key = packed_row.tobytes()
        st = self._state.get(key)
        if st is None:
            if len(self._state) >= self.LIMIT:
                self._state.clear()
            st = self._state[key] = self.b.unpack_state(packed_row)[0]
            self._packed.setdefault(st, packed_row.copy())
        return st

    def obs(self, obs_row: np.ndarray) -> tuple:
        """``obs_row``: one env's packed observations [A, ...] -> tuple of per-agent observations."""
        key = obs_row.tobytes()
        ob = self._obs.get(key)
        if ob is None:
            if len(self._obs) >= self.LIMIT:
                self._obs.clear()
            ob = self._obs[key] = self.b.unpack_obs(obs_row[None])[0]
        return ob


def to_numpy(t) -> np.ndarray:
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def trace_hash(hashes, out=None, init: bool = False, particle_offset: int = 0):
    """Folds one step's outputs (a ``StepOut`` of any model, with actions) into the per-particle rolling
    hash ``hashes`` (int64 [n], device) -- ``posg_trace_hash`` (include/posg_b200.h).  ``out=None`` with
    ``init=True`` only initialises."""
    n = hashes.shape[0]
    if out is None:
        args = (None, 0, None, 0, None, 0, None, None)
    else:
        args = (_lib.ptr(out.state), out.state.shape[1], _lib.ptr(out.obs), out.obs.numel() // max(n, 1),
                _lib.ptr(out.rewards), out.rewards.shape[1], _lib.ptr(out.flags), _lib.ptr(out.actions))
    with torch.cuda.device(hashes.device):
        _lib.check(_lib.lib().posg_trace_hash(n, _lib.ptr(hashes), int(init), particle_offset, *args, _lib.stream_ptr()))
    return hashes


class BatchedModel:
    """Subclasses set ``num_agents``, ``state_words``, ``obs_words``, ``noise_words``,
    ``num_picks`` and create ``self._handle``."""

    num_agents: int
    state_words: int
    obs_words: int
    noise_words: int
    num_picks: int

    def _init_device(self, device):
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
        if device is None:
            raise _lib.PosgB200Error("no CUDA device: posggym_baselines_b200 has no CPU fallback")
        self.device = torch.device(device)
        self._handle = C.c_void_p()

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and h.value:
            try:
                _lib.lib().posg_model_destroy(h)
            except Exception:
                pass
            self._handle = None

    def _empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def sample_initial_state(self, n: int, seed: int = 0, particle_offset: int = 0, picks=None, out=None):
        state = out if out is not None else self._empty((n, self.state_words), torch.int32)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().posg_model_sample_initial_state(
                self._handle, n, _lib.ptr(picks), seed, particle_offset, _lib.ptr(state), _lib.stream_ptr()))
        return state

    def observe(self, state, out=None):
        n = state.shape[0]
        obs = out if out is not None else self._empty((n, self.num_agents, self.obs_words), torch.int64)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().posg_model_observe(self._handle, n, _lib.ptr(state), _lib.ptr(obs), _lib.stream_ptr()))
        return obs

    def step(self, state, actions=None, noise=None, seed: int = 0, step_idx: int = 0, particle_offset: int = 0,
             out: Optional[StepOut] = None, live_count=None, want_actions: bool = True) -> StepOut:
        n, a = state.shape[0], self.num_agents
        if out is None:
            out = StepOut(
                self._empty((n, self.state_words), torch.int32), self._empty((n, a, self.obs_words), torch.int64),
                self._empty((n, a), torch.float64), self._empty((n,), torch.uint8),
                self._empty((n, a), torch.uint8) if (actions is None and want_actions) else None)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().posg_model_step(
                self._handle, n, _lib.ptr(state), _lib.ptr(actions), _lib.ptr(noise), seed, step_idx,
                particle_offset, _lib.ptr(out.state), _lib.ptr(out.obs), _lib.ptr(out.rewards),
                _lib.ptr(out.flags), _lib.ptr(out.actions), _lib.ptr(live_count), _lib.stream_ptr()))
        return out

    def step_host(self, state: np.ndarray, actions: np.ndarray, noise: np.ndarray):
        """``model.step`` for a few rows in host memory (``posg_model_step_host``): numpy in, numpy out --
        ``(state [n, state_words] int32, obs [n, A, obs_words] int64, rewards [n, A] float64, flags [n] uint8)``.
        The scalar facade and the planners' plugin path step through this."""
        n, a = state.shape[0], self.num_agents
        state = np.ascontiguousarray(state, dtype=np.int32)
        actions = np.ascontiguousarray(actions, dtype=np.uint8)
        noise = np.ascontiguousarray(noise, dtype=np.uint32).reshape(n, self.noise_words)
        out = (np.empty((n, self.state_words), np.int32), np.empty((n, a, self.obs_words), np.int64),
               np.empty((n, a), np.float64), np.empty((n,), np.uint8))
        _lib.check(_lib.lib().posg_model_step_host(self._handle, n, _lib.ptr(state), _lib.ptr(actions), _lib.ptr(noise),
                                                   4 * self.noise_words, *(_lib.ptr(o) for o in out)))
        return out

    def rollout(self, state, ego_agent: int, max_steps: int, discount: float, seed: int = 0, step0: int = 0,
                particle_offset: int = 0, want_state: bool = False, live_count=None):
        n = state.shape[0]
        returns = self._empty((n,), torch.float64)
        steps = self._empty((n,), torch.uint8)
        final = self._empty((n, self.state_words), torch.int32) if want_state else None
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().posg_model_rollout(
                self._handle, n, _lib.ptr(state), ego_agent, max_steps, discount, seed, step0, particle_offset,
                _lib.ptr(returns), _lib.ptr(steps), _lib.ptr(final), _lib.ptr(live_count), _lib.stream_ptr()))
        return returns, steps, final


class ModelFacade:
    """Scalar ``posggym.model.POSGModel`` surface over a BatchedModel (batch of one)."""

    is_symmetric = True

    def _init_facade(self, batched: BatchedModel, env_id: str, max_episode_steps: int):
        self.batched = batched
        self.possible_agents = tuple(str(i) for i in range(batched.num_agents))
        self.spec = Spec(env_id, max_episode_steps)
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
        # the picks come from model.rng one by one; the state is built by the kernel from the explicit picks
        b = self.batched
        picks: List[int] = []
        for _ in range(b.num_picks):
            picks.append(self._rng.randrange(b.initial_pick_bounds(picks)))
        st = b.sample_initial_state(1, picks=torch.tensor([picks], dtype=torch.uint8, device=b.device))
        return b.unpack_state(st)[0]

    def sample_initial_obs(self, state):
        b = self.batched
        packed = torch.from_numpy(b.pack_state([state])).to(b.device)
        obs = b.unpack_obs(b.observe(packed))[0]
        return {i: obs[k] for k, i in enumerate(self.possible_agents)}

    def sample_agent_initial_state(self, agent_id: str, obs):
        raise NotImplementedError

    def host_step_noise(self, n: int) -> np.ndarray:
        """Explicit noise for ``n`` rows of ``batched.step_host`` / ``batched.step``: what ``step`` draws from
        ``self.rng``, n times (uint32 [n, noise_words])."""
        b = self.batched
        if getattr(b, "noise_is_dummy", False):     # deterministic transition: nothing is drawn
            return np.zeros((n, b.noise_words), np.uint32)
        return np.array([[self._rng.getrandbits(32) for _ in range(b.noise_words)] for _ in range(n)],
                        dtype=np.uint32).reshape(n, b.noise_words)

    def batched_step_noise(self, n: int):
        return torch.from_numpy(self.host_step_noise(n).view(np.int32)).to(self.batched.device)

    def step(self, state, actions: Dict[str, int]) -> JointTimestep:
        b = self.batched
        act = np.array([[int(actions[i]) for i in self.possible_agents]], dtype=np.uint8)
        codec = self.__dict__.get("_codec") or self.__dict__.setdefault("_codec", ScalarCodec(b))
        st, ob, rw, fl = b.step_host(codec.pack(state), act, self.host_step_noise(1))
        nstate = codec.state(st)
        obs = codec.obs(ob[0])
        rew = rw[0]
        flags = int(fl[0])
        ids = self.possible_agents
        return JointTimestep(
            nstate, {i: obs[k] for k, i in enumerate(ids)}, {i: float(rew[k]) for k, i in enumerate(ids)},
            {i: bool((flags >> k) & 1) for k, i in enumerate(ids)}, {i: False for i in ids},
            bool(flags & FLAG_ALL_DONE), {i: {} for i in ids})
