"""A CPU stand-in for the device model facade, for tests of the planners' HOST logic (planning/host_search.py,
planning/potmmcp.py, planning/host_intmcp.py): the batch interface those engines call (``pack_state`` /
``unpack_state`` / ``pack_obs`` / ``unpack_obs`` / ``observe`` / ``step_host``) answered by the CPU oracle model, with
states and observations "packed" as indices into a registry.  Test infrastructure only -- the product planners get
their facade from posggym_baselines_b200.envs and fail without the CUDA library."""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch

from oracle.driving import DrivingModel


class _Registry:
    def __init__(self):
        self.items, self.index = [], {}

    def put(self, x) -> int:
        k = self.index.get(x)
        if k is None:
            k = self.index[x] = len(self.items)
            self.items.append(x)
        return k


class RegistryBatched:
    device = torch.device("cpu")

    def __init__(self, model):
        self.m, self.states, self.obs = model, _Registry(), _Registry()
        self.ids = list(model.possible_agents)

    def pack_state(self, states):
        return np.array([[self.states.put(s)] for s in states], dtype=np.int64)

    def unpack_state(self, t):
        return [self.states.items[int(k)] for k in t.reshape(-1)]

    def pack_obs(self, obs):
        return np.array([self.obs.put(obs)], dtype=np.uint64)

    def unpack_obs(self, t):
        return [tuple(self.obs.items[int(k)] for k in row.reshape(-1)) for row in t]

    def _obs_rows(self, joint_obs_list):
        return torch.tensor([[[self.obs.put(jo[i])] for i in self.ids] for jo in joint_obs_list], dtype=torch.int64)

    def observe(self, states):
        return self._obs_rows([self.m.sample_initial_obs(s) for s in self.unpack_state(states)])

    def step_host(self, st, act, noise=None):
        rows = [self.m.step(s, {i: int(a[k]) for k, i in enumerate(self.ids)}) for s, a in zip(self.unpack_state(st), act)]
        flags = [sum(int(ts.terminations[i]) << k for k, i in enumerate(self.ids)) | (0x80 if ts.all_done else 0) for ts in rows]
        return (self.pack_state([ts.state for ts in rows]), self._obs_rows([ts.observations for ts in rows]).numpy(),
                np.array([[ts.rewards[i] for i in self.ids] for ts in rows], dtype=np.float64), np.array(flags, dtype=np.uint8))


    def rollout(self, st, ego_agent, max_steps, discount, seed=0, step0=0, particle_offset=0):
        """The fused uniform-random rollout kernel's contract: discounted ego return, stops after the step on which the
        ego agent terminated or everything is done; (returns, steps executed, None)."""
        import random as _random

        rng = _random.Random((seed << 20) ^ particle_offset)
        rets, steps = [], []
        for s in self.unpack_state(st):
            ret, k = 0.0, 0
            while k < max_steps:
                ts = self.m.step(s, {i: rng.randrange(self.m.action_spaces[i].n) for i in self.ids})
                ret += discount**k * ts.rewards[self.ids[ego_agent]]
                k += 1
                if ts.terminations[self.ids[ego_agent]] or ts.all_done:
                    break
                s = ts.state
            rets.append(ret)
            steps.append(k)
        return torch.tensor(rets, dtype=torch.float64), torch.tensor(steps, dtype=torch.uint8), None


class DoubleDrivingModel(DrivingModel):
    """The oracle Driving model with a ``.batched`` the host engines can call."""

    def __init__(self, **kw):
        super().__init__(**kw)
        if self.spec is None:
            self.spec = SimpleNamespace(max_episode_steps=50, id="Driving-v1")
        self.batched = RegistryBatched(self)

    def host_step_noise(self, n):
        return None

    def initial_belief(self, agent_id, init_obs, need):
        """The rejection loop the device runs in ``posg_belief_initial``."""
        out = []
        while len(out) < need:
            s = self.sample_initial_state()
            if self.sample_initial_obs(s)[agent_id] == init_obs:
                out.append(s)
        return torch.from_numpy(self.batched.pack_state(out))
