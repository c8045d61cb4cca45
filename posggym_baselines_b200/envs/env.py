"""Scalar environment loop over a model facade -- the surface of ``posggym.Env`` the
reference's README / test loops use (README.md:82-88, tests/planning/test_pomcp.py:16-33):
``reset() -> (obs, infos)``, ``step(actions) -> (obs, rewards, terminations, truncations,
all_done, infos)``, ``model``, ``possible_agents``; the 50-step TimeLimit of the
registered envs is applied here (truncation is not the model's business)."""
from __future__ import annotations

from typing import Dict, Optional


class Env:
    def __init__(self, model, max_episode_steps: Optional[int] = None):
        self.model = model
        self.max_episode_steps = max_episode_steps if max_episode_steps is not None else (
            model.spec.max_episode_steps if model.spec is not None else None)
        self._state = None
        self._t = 0

    @property
    def possible_agents(self):
        return self.model.possible_agents

    @property
    def action_spaces(self):
        return self.model.action_spaces

    @property
    def observation_spaces(self):
        return self.model.observation_spaces

    @property
    def state(self):
        return self._state

    def reset(self, seed: Optional[int] = None, options=None):
        if seed is not None:
            self.model.seed(seed)
        self._state = self.model.sample_initial_state()
        self._t = 0
        return self.model.sample_initial_obs(self._state), {i: {} for i in self.possible_agents}

    def step(self, actions: Dict[str, int]):
        ts = self.model.step(self._state, actions)
        self._state = ts.state
        self._t += 1
        trunc, all_done = dict(ts.truncations), ts.all_done
        if self.max_episode_steps is not None and self._t >= self.max_episode_steps:
            trunc = {i: True for i in trunc}
            all_done = True
        return ts.observations, ts.rewards, ts.terminations, trunc, all_done, ts.infos

    def render(self):
        return None

    def close(self):
        pass


def make(env_id: str, **kwargs) -> Env:
    """``posggym.make`` for the in-scope envs (GPU model facades)."""
    kwargs.pop("render_mode", None)
    max_steps = kwargs.pop("max_episode_steps", 50)
    if env_id == "Driving-v1":
        from posggym_baselines_b200.envs.driving import DrivingModel

        return Env(DrivingModel(**kwargs), max_steps)
    if env_id == "PredatorPrey-v0":
        from posggym_baselines_b200.envs.predator_prey import PredatorPreyModel

        return Env(PredatorPreyModel(**kwargs), max_steps)
    if env_id == "LevelBasedForaging-v3":
        from posggym_baselines_b200.envs.lbf import LBFModel

        return Env(LBFModel(**kwargs), max_steps)
    raise KeyError(f"{env_id!r} is not available in posggym_baselines_b200")
