"""GPU vector env with the contract of the reference's
``StackEnv(RecordEpisodeStatistics(SyncVectorEnv([FlattenObservations(env)] * n)))``
(/root/reference/posggym_baselines/ppo/config.py:223-232) as the rollout worker uses it
(ppo/worker.py:82-86,176-216):

    obs, infos = envs.reset()
    obs, reward, terminated, truncated, dones, infos = envs.step(flat_int_actions)

``obs`` has ``num_envs * num_agents`` rows (env-major) of one-hot flattened float32
observations, ``reward`` / ``terminated`` / ``truncated`` the same number of rows, ``dones``
one entry per env, ``infos[agent_id]["episode"]`` holds ``r``, ``l``, ``t`` arrays indexed by
env on the steps where some episode ended.  Finished envs are auto-reset in the same call
(the worker never calls ``reset`` again).  Everything is computed by one kernel launch
(posg_vec_step) without a host synchronisation; tensors stay on the GPU unless ``as_numpy=True``
and are views of buffers that the next call overwrites (the worker copies them into its rollout
buffers right away, ppo/worker.py:141-146,176-195).

Sharding: ``env_offset`` gives this shard's first global env index (Philox counters are
global env indices, so a 65 536-env job split over 8 GPUs is the same job).
"""
from __future__ import annotations

import time
from typing import Dict, Optional

import numpy as np
import torch

from posggym_baselines_b200 import _lib


class BatchedVecEnv:
    def __init__(self, model, num_envs: int, max_episode_steps: Optional[int] = None, seed: int = 0,
                 env_offset: int = 0, as_numpy: bool = False, write_obs: bool = True):
        self.model = model
        self._b = model.batched
        self.num_envs = int(num_envs)
        self.possible_agents = tuple(model.possible_agents)
        self.num_agents = len(self.possible_agents)
        self.max_episode_steps = int(max_episode_steps if max_episode_steps is not None
                                     else model.spec.max_episode_steps)
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.env_offset = int(env_offset)
        self.as_numpy = as_numpy
        self.action_spaces = model.action_spaces
        dev = self._b.device
        n, a = self.num_envs, self.num_agents
        lib = _lib.lib()
        self.flat_dim = _lib.check(lib.posg_vec_flat_dim(self._b._handle))
        sw = getattr(self._b, "state_words", 2)
        self._state = torch.zeros((n, sw), dtype=torch.int32, device=dev)
        self._ep_t = torch.zeros(n, dtype=torch.int16, device=dev)
        self._ep_ret = torch.zeros((n, a), dtype=torch.float64, device=dev)
        self._ep_count = torch.zeros(n, dtype=torch.int32, device=dev)
        # write_obs=False: no one-hot rows (the policy consumes the packed observation through embed())
        self.write_obs = bool(write_obs)
        self._obs = torch.empty((n * a, self.flat_dim), dtype=torch.float32, device=dev) if write_obs else None
        self._wT, self._wT_key = None, None
        self._rew = torch.zeros(n * a, dtype=torch.float32, device=dev)
        self._term = torch.zeros(n * a, dtype=torch.uint8, device=dev)
        self._trunc = torch.zeros(n * a, dtype=torch.uint8, device=dev)
        self._done = torch.zeros(n, dtype=torch.uint8, device=dev)
        # bool views of the kernel's byte flags (the kernel writes exactly 0 / 1)
        self._term_b, self._trunc_b, self._done_b = (
            self._term.view(torch.bool), self._trunc.view(torch.bool), self._done.view(torch.bool))
        self._stat_r = torch.zeros((n, a), dtype=torch.float32, device=dev)
        self._stat_l = torch.zeros(n, dtype=torch.int32, device=dev)
        self._t0 = torch.zeros(n, dtype=torch.float64, device=dev)      # start time of the running episode
        self._stat_t = torch.zeros(n, dtype=torch.float64, device=dev)  # wall time of the episode just ended
        self._global_step = 0
        self._step_ctr = torch.zeros(2, dtype=torch.int32, device=dev)   # [step index, CTA scratch] for captured graphs
        self._device_counter = False
        self._args = None
        self._ret = None
        self._dev_index = dev.index if dev.index is not None else torch.cuda.current_device()

    @property
    def single_observation_shape(self):
        return (self.flat_dim,)

    def _launch(self, actions, reset_only: int, graphed: bool = False, outs=None):
        # hot loop of the collector: buffer pointers are cached (the buffers live as long as the
        # env), the device guard is only entered when another device is current
        if self._args is None:
            self._fn = _lib.lib().posg_vec_step
            self._args = (
                (self._b._handle, self.num_envs, _lib.ptr(self._state), _lib.ptr(self._ep_t), _lib.ptr(self._ep_ret),
                 _lib.ptr(self._ep_count)),
                (_lib.ptr(self._obs) if self.write_obs else None, _lib.ptr(self._rew), _lib.ptr(self._term),
                 _lib.ptr(self._trunc),
                 _lib.ptr(self._done), _lib.ptr(self._stat_r), _lib.ptr(self._stat_l)),
                (_lib.ptr(self._t0), _lib.ptr(self._stat_t)))
        head, tail, clock = self._args
        if outs is not None:
            # the collector's rollout storage: rewards / terminated / truncated of this step go straight into its
            # [T, rows] buffers (no copy kernels between the launches of a captured rollout)
            rew, term, trunc = outs
            tail = (tail[0], _lib.ptr(rew), _lib.ptr(term), _lib.ptr(trunc)) + tail[4:]
        a = None if actions is None else actions.data_ptr()
        now = time.perf_counter()
        # device-counter mode (capture_step / enable_device_counter): the step index lives on the
        # device for eager and captured launches alike; a captured launch has no wall clock
        ctr = self._step_ctr.data_ptr() if self._device_counter else None
        if graphed:
            clock = (None, None)
        if torch.cuda.current_device() == self._dev_index:
            rc = self._fn(*head, a, self.seed, self._global_step, self.env_offset, self.max_episode_steps, reset_only,
                          *tail, now, *clock, ctr, torch.cuda.current_stream().cuda_stream or None)
        else:
            with torch.cuda.device(self._b.device):
                rc = self._fn(*head, a, self.seed, self._global_step, self.env_offset, self.max_episode_steps,
                              reset_only, *tail, now, *clock, ctr, torch.cuda.current_stream().cuda_stream or None)
        if rc < 0:
            _lib.check(rc)

    def _out(self, t):
        if t is None:    # write_obs=False
            return None
        return t.cpu().numpy() if self.as_numpy else t

    def reset(self, seed: Optional[int] = None, options=None):
        if seed is not None:
            self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self._global_step = 0
        if self._device_counter:
            self._step_ctr.zero_()
        self._launch(None, 1)
        return self._out(self._obs), {i: {} for i in self.possible_agents}

    def step(self, actions, _replay_only: bool = False):
        """``actions``: flat integer array / tensor of ``num_envs * num_agents`` entries, env-major.

        One kernel launch and no host synchronisation: ``infos[agent]["episode"]`` is present on
        every step (the worker only reads it at the envs whose ``dones`` entry is set,
        ppo/worker.py:196-216); ``r`` / ``l`` / ``t`` are zero at the other envs, like the arrays
        ``RecordEpisodeStatistics`` builds on the steps where some episode ended."""
        if _replay_only:   # capture_step: just the tuple of persistent views
            if self.as_numpy:
                raise _lib.PosgB200Error("capture_step needs as_numpy=False")
            infos = {i: {"episode": {"r": self._stat_r[:, k], "l": self._stat_l, "t": self._stat_t}}
                     for k, i in enumerate(self.possible_agents)}
            return (self._obs, self._rew, self._term_b, self._trunc_b, self._done_b, infos)
        if not isinstance(actions, torch.Tensor):
            actions = torch.as_tensor(np.asarray(actions))
        if actions.dtype != torch.int64 or actions.device != self._b.device or not actions.is_contiguous():
            actions = actions.to(device=self._b.device, dtype=torch.int64).contiguous()
        assert actions.numel() == self.num_envs * self.num_agents
        self._global_step += 1
        self._launch(actions, 0)
        if not self.as_numpy:
            if self._ret is None:   # views of the persistent buffers: built once
                infos = {i: {"episode": {"r": self._stat_r[:, k], "l": self._stat_l, "t": self._stat_t}}
                         for k, i in enumerate(self.possible_agents)}
                self._ret = (self._obs, self._rew, self._term_b, self._trunc_b, self._done_b, infos)
            return self._ret
        infos: Dict[str, dict] = {}
        for k, i in enumerate(self.possible_agents):
            infos[i] = {"episode": {"r": self._out(self._stat_r[:, k]), "l": self._out(self._stat_l),
                                    "t": self._out(self._stat_t)}}
        return (self._out(self._obs), self._out(self._rew), self._out(self._term_b), self._out(self._trunc_b),
                self._out(self._done_b), infos)

    def enable_device_counter(self):
        """Keep the step index of the Philox counter in device memory (advanced by the kernel) from
        now on, for eager and captured launches alike -- the precondition for capturing
        ``step_device`` in a CUDA graph, where scalar arguments are frozen."""
        if not self._device_counter:
            self._step_ctr.zero_()
            self._step_ctr[0] = self._global_step
            self._device_counter = True

    def step_device(self, actions: torch.Tensor, outs=None):
        """Graph-capturable ``envs.step``: ``actions`` must be an int64 CUDA tensor of
        ``num_envs * num_agents`` entries; no host-side conversion, clock or synchronisation.  Call
        it inside ``torch.cuda.graph(...)`` (after ``enable_device_counter()``) together with the
        policy's kernels to replay a whole rollout with one launch.  Returns the same tuple of
        persistent views as ``step``; episode wall times are not tracked in captured launches."""
        capturing = torch.cuda.is_current_stream_capturing()
        if not capturing:
            self.enable_device_counter()
        assert self._device_counter, "call enable_device_counter() before capturing"
        self._global_step += 1    # informational in this mode: the device counter is authoritative
        self._launch(actions, 0, graphed=capturing, outs=outs)
        return self.step(None, _replay_only=True)

    def step_embed_device(self, actions: torch.Tensor, w_t: torch.Tensor, bias: torch.Tensor, x_out: torch.Tensor,
                          outs=None, state_copy: Optional[torch.Tensor] = None, tanh: bool = True):
        """``step_device`` FUSED with ``embed`` of the observations the step returns (posg_vec_step_embed): one
        launch steps the envs and writes ``x_out`` = trunk layer 0 of the policy for the NEXT step (and the new
        packed state into ``state_copy``).  ``w_t`` = ``trunk.0.weight`` transposed ``[obs_dim, 64]``.  ``outs`` =
        ``(rewards, terminated, truncated)`` rows of the rollout storage (default: the env's own buffers)."""
        capturing = torch.cuda.is_current_stream_capturing()
        if not capturing:
            self.enable_device_counter()
        assert self._device_counter, "call enable_device_counter() before capturing"
        self._global_step += 1
        rew, term, trunc = outs if outs is not None else (self._rew, self._term, self._trunc)
        with torch.cuda.device(self._b.device):
            _lib.check(_lib.lib().posg_vec_step_embed(
                self._b._handle, self.num_envs, _lib.ptr(self._state), _lib.ptr(self._ep_t), _lib.ptr(self._ep_ret),
                _lib.ptr(self._ep_count), actions.data_ptr(), self.seed, self._global_step, self.env_offset,
                self.max_episode_steps, _lib.ptr(rew), _lib.ptr(term), _lib.ptr(trunc), _lib.ptr(self._done),
                _lib.ptr(self._stat_r), _lib.ptr(self._stat_l), self._step_ctr.data_ptr(), _lib.ptr(w_t), _lib.ptr(bias),
                int(w_t.shape[1]), int(tanh), _lib.ptr(x_out), None if state_copy is None else _lib.ptr(state_copy),
                _lib.stream_ptr()))

    def capture_step(self):
        """Capture one ``envs.step`` in a CUDA graph.  Returns ``(actions, replay)``: write the step's
        actions into the static int64 tensor ``actions``, call ``replay()`` -- it returns the same
        tuple as ``step`` (views of the same persistent buffers)."""
        dev = self._b.device
        actions = torch.zeros(self.num_envs * self.num_agents, dtype=torch.int64, device=dev)
        self.enable_device_counter()
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):      # recorded, not executed: the envs do not move here
            ret = self.step_device(actions)
        self._global_step -= 1             # the recorded launch was not a step

        def replay():
            self._global_step += 1
            graph.replay()
            return ret

        self._graph = graph   # keep alive
        return actions, replay

    def embed(self, weight, bias, tanh: bool = True, out=None):
        """First trunk layer of the PPO policy on the CURRENT observations without the one-hot rows:
        ``tanh(F.linear(obs, weight, bias))`` for ``weight`` = ``trunk.0.weight`` ``[hidden, obs_dim]``
        and ``bias`` = ``trunk.0.bias`` of ``PPOLSTMModel`` / ``PPOMLPModel`` (ppo/network.py:171-178),
        computed as a sum of one weight column per observation field (posg_vec_embed).  Returns
        ``[num_envs * num_agents, hidden]`` float32.  Construct the env with ``write_obs=False`` to
        make ``reset`` / ``step`` skip the one-hot rows altogether."""
        key = (weight.data_ptr(), weight._version, tuple(weight.shape), bias.data_ptr(), bias._version)
        if self._wT_key != key:   # transposed copy: one contiguous row of `hidden` floats per one-hot input
            assert weight.shape[1] == self.flat_dim, (tuple(weight.shape), self.flat_dim)
            self._wT = weight.detach().to(device=self._b.device, dtype=torch.float32).t().contiguous()
            self._wT_key = key
            self._bias = bias.detach().to(device=self._b.device, dtype=torch.float32).contiguous().clone()
        b = self._bias
        h = self._wT.shape[1]
        if out is None:
            out = torch.empty((self.num_envs * self.num_agents, h), dtype=torch.float32, device=self._b.device)
        with torch.cuda.device(self._b.device):
            _lib.check(_lib.lib().posg_vec_embed(
                self._b._handle, self.num_envs, _lib.ptr(self._state), _lib.ptr(self._wT), _lib.ptr(b), h, int(tanh),
                _lib.ptr(out), _lib.stream_ptr()))
        return out

    def close(self):
        pass


def load_vec_env(env_id: str, env_kwargs: Optional[dict] = None, num_envs: int = 32, seed: int = 0,
                 env_offset: int = 0, device=None, as_numpy: bool = False) -> BatchedVecEnv:
    """``PPOConfig.load_vec_env`` (ppo/config.py:223-232) for the in-scope envs."""
    from posggym_baselines_b200.envs.env import make

    kw = dict(env_kwargs or {})
    if device is not None:
        kw["device"] = device
    env = make(env_id, **kw)
    return BatchedVecEnv(env.model, num_envs, env.max_episode_steps, seed=seed, env_offset=env_offset,
                         as_numpy=as_numpy)
