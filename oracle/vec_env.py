"""Vector-env stack of the PPO collector (oracle; test infrastructure only).

Restates the behaviour the reference's rollout worker relies on
(/root/reference/posggym_baselines/ppo/config.py:223-232 builds
``StackEnv(RecordEpisodeStatistics(SyncVectorEnv([thunk] * num_envs)))`` and
ppo/worker.py:82-86,176-216 consumes it) of the un-vendored posggym modules
``posggym/vector/sync_vector_env.py`` and ``posggym/wrappers/{flatten_observations,
record_episode_statistics,stack_env}.py``:

* ``SyncVectorEnv`` steps its sub-envs serially and auto-resets a sub-env on the step
  its ``all_done`` fires, returning the reset observation (the worker never calls
  ``reset`` again: worker.py:83 is the only call);
* ``RecordEpisodeStatistics`` adds ``infos[agent]["episode"][k]`` arrays (k in r, l, t)
  indexed by sub-env, filled on the step an episode ends (worker.py:208-216);
* ``StackEnv`` stacks the per-agent dicts env-major: row ``env * num_agents + agent``
  (the worker reshapes to ``(num_envs, num_agents, ...)``: worker.py:82-86,181-195);
  actions arrive as a flat integer array of the same row order (worker.py:176-178);
* ``FlattenObservations`` one-hot flattens each agent's observation (gymnasium
  ``spaces.flatten``), dtype of the flattened Box (int64 for Discrete trees).
"""
from __future__ import annotations

import time
from typing import Callable, Dict, List, Optional

import numpy as np


class FlattenObservations:
    def __init__(self, env):
        from gymnasium import spaces

        self.env = env
        self.model = env.model
        self._spaces = spaces
        self._src = env.observation_spaces
        self.observation_spaces = {i: spaces.flatten_space(s) for i, s in self._src.items()}
        self.action_spaces = env.action_spaces
        self.possible_agents = env.possible_agents

    def _flat(self, obs):
        return {i: self._spaces.flatten(self._src[i], o) for i, o in obs.items()}

    def reset(self, seed=None, options=None):
        obs, info = self.env.reset(seed=seed, options=options)
        return self._flat(obs), info

    def step(self, actions):
        obs, rew, term, trunc, done, info = self.env.step(actions)
        return self._flat(obs), rew, term, trunc, done, info

    def close(self):
        self.env.close()


class SyncVectorEnv:
    def __init__(self, env_fns: List[Callable]):
        self.envs = [fn() for fn in env_fns]
        self.num_envs = len(self.envs)
        e0 = self.envs[0]
        self.possible_agents = e0.possible_agents
        self.single_observation_spaces = e0.observation_spaces
        self.single_action_spaces = e0.action_spaces
        self.observation_spaces = e0.observation_spaces
        self.action_spaces = e0.action_spaces

    def reset(self, seed=None, options=None):
        obs = {i: [] for i in self.possible_agents}
        for k, e in enumerate(self.envs):
            o, _ = e.reset(seed=None if seed is None else seed + k)
            for i in self.possible_agents:
                obs[i].append(o[i])
        return {i: np.stack(v) for i, v in obs.items()}, {i: {} for i in self.possible_agents}

    def step(self, actions: Dict[str, np.ndarray]):
        A = self.possible_agents
        obs = {i: [] for i in A}
        rew = {i: np.zeros(self.num_envs, np.float64) for i in A}
        term = {i: np.zeros(self.num_envs, bool) for i in A}
        trunc = {i: np.zeros(self.num_envs, bool) for i in A}
        dones = np.zeros(self.num_envs, bool)
        for k, e in enumerate(self.envs):
            o, r, te, tr, d, _ = e.step({i: int(actions[i][k]) for i in A})
            if d:
                o, _ = e.reset()
            for i in A:
                obs[i].append(o[i])
                rew[i][k], term[i][k], trunc[i][k] = r[i], te[i], tr[i]
            dones[k] = d
        return {i: np.stack(v) for i, v in obs.items()}, rew, term, trunc, dones, {i: {} for i in A}

    def close(self):
        for e in self.envs:
            e.close()


class RecordEpisodeStatistics:
    def __init__(self, env: SyncVectorEnv):
        self.env = env
        self.num_envs = env.num_envs
        self.possible_agents = env.possible_agents
        self.observation_spaces = env.observation_spaces
        self.action_spaces = env.action_spaces
        self._ret = {i: np.zeros(self.num_envs) for i in self.possible_agents}
        self._len = np.zeros(self.num_envs, np.int64)
        self._t0 = np.full(self.num_envs, time.perf_counter())

    def reset(self, seed=None, options=None):
        out = self.env.reset(seed=seed, options=options)
        for i in self.possible_agents:
            self._ret[i][:] = 0
        self._len[:] = 0
        self._t0[:] = time.perf_counter()
        return out

    def step(self, actions):
        obs, rew, term, trunc, dones, infos = self.env.step(actions)
        self._len += 1
        for i in self.possible_agents:
            self._ret[i] += rew[i]
        if dones.any():
            now = time.perf_counter()
            for i in self.possible_agents:
                infos[i]["episode"] = {
                    "r": np.where(dones, self._ret[i], 0.0),
                    "l": np.where(dones, self._len, 0),
                    "t": np.where(dones, now - self._t0, 0.0),
                }
                self._ret[i][dones] = 0
            self._len[dones] = 0
            self._t0[dones] = now
        return obs, rew, term, trunc, dones, infos

    def close(self):
        self.env.close()


class StackEnv:
    def __init__(self, env):
        self.env = env
        self.num_envs = env.num_envs
        self.possible_agents = env.possible_agents
        self.num_agents = len(self.possible_agents)
        self.observation_spaces = env.observation_spaces
        self.action_spaces = env.action_spaces

    def _stack(self, d):
        # (num_agents, num_envs, ...) -> env-major rows
        a = np.stack([d[i] for i in self.possible_agents], axis=1)
        return a.reshape((self.num_envs * self.num_agents,) + a.shape[2:])

    def reset(self, seed=None, options=None):
        obs, infos = self.env.reset(seed=seed, options=options)
        return self._stack(obs), infos

    def step(self, actions):
        a = np.asarray(actions).reshape(self.num_envs, self.num_agents)
        act = {i: a[:, k] for k, i in enumerate(self.possible_agents)}
        obs, rew, term, trunc, dones, infos = self.env.step(act)
        return (self._stack(obs), self._stack(rew).astype(np.float32), self._stack(term), self._stack(trunc),
                dones, infos)

    def close(self):
        self.env.close()
