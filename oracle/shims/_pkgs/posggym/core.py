"""posggym.Env / posggym.make stand-in (oracle shim): DefaultEnv + TimeLimit."""
from __future__ import annotations

from typing import Dict, Optional

from posggym.model import _Spec

_MAX_STEPS = {"Driving-v1": 50, "PredatorPrey-v0": 50, "LevelBasedForaging-v3": 50}


class Env:
    def __init__(self, model, max_episode_steps: Optional[int]):
        self.model = model
        self.max_episode_steps = max_episode_steps
        self._state = None
        self._t = 0

    @property
    def possible_agents(self):
        return self.model.possible_agents

    @property
    def agents(self):
        return list(self.model.possible_agents)

    @property
    def action_spaces(self):
        return self.model.action_spaces

    @property
    def observation_spaces(self):
        return self.model.observation_spaces

    @property
    def state(self):
        return self._state

    @property
    def unwrapped(self):
        return self

    def reset(self, seed: Optional[int] = None, options=None):
        if seed is not None:
            self.model.seed(seed)
        self._state = self.model.sample_initial_state()
        self._t = 0
        obs = self.model.sample_initial_obs(self._state)
        return obs, {i: {} for i in self.possible_agents}

    def step(self, actions: Dict[str, int]):
        ts = self.model.step(self._state, actions)
        self._state = ts.state
        self._t += 1
        trunc = dict(ts.truncations)
        all_done = ts.all_done
        if self.max_episode_steps is not None and self._t >= self.max_episode_steps:
            trunc = {i: True for i in trunc}
            all_done = True
        return ts.observations, ts.rewards, ts.terminations, trunc, all_done, ts.infos

    def render(self):
        return None

    def close(self):
        pass


def make(env_id: str, **kwargs) -> Env:
    kwargs.pop("render_mode", None)
    max_steps = kwargs.pop("max_episode_steps", _MAX_STEPS.get(env_id))
    if env_id == "Driving-v1":
        from oracle.driving import DrivingModel as cls
    elif env_id == "PredatorPrey-v0":
        from oracle.predator_prey import PredatorPreyModel as cls
    elif env_id == "LevelBasedForaging-v3":
        from oracle.lbf import LevelBasedForagingModel as cls
    else:
        raise KeyError(env_id)
    model = cls(**kwargs)
    model.spec = _Spec(env_id, max_steps)
    return Env(model, max_steps)
