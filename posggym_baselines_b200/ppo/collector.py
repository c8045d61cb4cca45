"""The rollout collector's inner loop on the device (ppo/worker.py:135-224 of the reference): per step
``policy.get_action_and_value`` on the current observations, ``envs.step`` with the sampled actions, store
reward / done / value / log-prob.  TWO launches per step -- ``posg_policy_step`` (the network after trunk layer 0 on
the tensor cores, samples the actions) and ``posg_vec_step_embed`` (env step, TimeLimit, statistics, auto-reset, AND
trunk layer 0 of the next step on the observations it returns: no one-hot rows, nothing recomputed; rewards,
terminated / truncated and the packed states go straight into the rollout storage) -- and the whole
``rollout_length``-step loop is captured in ONE CUDA graph: the step index of every Philox counter lives in
device memory, so a replay is a fresh rollout."""
from __future__ import annotations

from typing import Dict, Optional

import torch

from posggym_baselines_b200.ppo.policy import FusedPolicy
from posggym_baselines_b200.vector.vec_env import BatchedVecEnv


class DeviceCollector:
    def __init__(self, model, state_dict: Dict[str, torch.Tensor], num_envs: int, rollout_length: int = 64, seed: int = 0,
                 env_offset: int = 0):
        self.envs = BatchedVecEnv(model, num_envs, seed=seed, env_offset=env_offset, write_obs=False)
        dev = self.envs._b.device
        self.policy = FusedPolicy(state_dict, device=dev)
        self.T, self.rows = int(rollout_length), num_envs * self.envs.num_agents
        self.seed, self.env_offset = seed, env_offset
        T, R = self.T, self.rows
        f32 = dict(dtype=torch.float32, device=dev)
        self.h = torch.zeros((R, 64), **f32)
        self.c = torch.zeros((R, 64), **f32)
        self.x0 = torch.empty((R, 64), **f32)
        # rollout storage [T, rows] (the worker's buffers, worker.py:96-133)
        self.actions = torch.zeros((T, R), dtype=torch.int64, device=dev)
        self.logprobs, self.values, self.entropies = (torch.zeros((T, R), **f32) for _ in range(3))
        self.rewards = torch.zeros((T, R), **f32)
        # terminated / truncated as envs.step returned them: row t + 1 is written by step t, row 0 carries over from
        # the previous rollout; dones = their OR, computed once per rollout
        self.terms = torch.zeros((T + 1, R), dtype=torch.uint8, device=dev)
        self.truncs = torch.zeros((T + 1, R), dtype=torch.uint8, device=dev)
        self.dones = torch.zeros((T + 1, R), dtype=torch.uint8, device=dev)
        self.states = torch.zeros((T + 1, num_envs, self.envs._state.shape[1]), dtype=torch.int32, device=dev)
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self.envs.reset()
        self.envs.enable_device_counter()
        self.states[T].copy_(self.envs._state)
        self._embed()          # x0 of the first step; afterwards every env step leaves the next x0 behind

    def _step(self, t: int):
        e, p = self.envs, self.policy
        out = {"action": self.actions[t], "logprob": self.logprobs[t], "entropy": self.entropies[t], "value": self.values[t]}
        p.step(self.x0, (self.h, self.c), None, terminated=self.terms[t], truncated=self.truncs[t], seed=self.seed,
               row_offset=self.env_offset * e.num_agents, step_device=e._step_ctr, out=out)
        # states[t + 1] = the packed observations' source of the next step (the learner re-embeds it)
        e.step_embed_device(self.actions[t], p.w0_t, p.b0, self.x0,
                            outs=(self.rewards[t], self.terms[t + 1], self.truncs[t + 1]), state_copy=self.states[t + 1])

    def _embed(self):
        from posggym_baselines_b200 import _lib

        e, p = self.envs, self.policy
        with torch.cuda.device(e._b.device):
            _lib.check(_lib.lib().posg_vec_embed(e._b._handle, e.num_envs, _lib.ptr(e._state), _lib.ptr(p.w0_t), _lib.ptr(p.b0),
                                                 64, 1, _lib.ptr(self.x0), _lib.stream_ptr()))

    def rollout_eager(self):
        self.terms[0].copy_(self.terms[self.T])      # next_done of the previous rollout (zeros after reset)
        self.truncs[0].copy_(self.truncs[self.T])
        self.states[0].copy_(self.states[self.T])
        for t in range(self.T):
            self._step(t)
        torch.bitwise_or(self.terms, self.truncs, out=self.dones)

    def capture(self):
        """Captures one whole rollout; afterwards ``rollout()`` is a single graph launch."""
        dev = self.envs._b.device
        self.rollout_eager()                      # warm-up (allocations, lazy inits)
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.rollout_eager()
        self._graph = g

    def rollout(self):
        if self._graph is None:
            self.rollout_eager()
        else:
            self._graph.replay()
        return {"actions": self.actions, "logprobs": self.logprobs, "values": self.values, "entropies": self.entropies,
                "rewards": self.rewards, "dones": self.dones, "states": self.states[:self.T]}
