"""posggym.agents.policy stand-in."""
from __future__ import annotations

from typing import Any, Dict

PolicyState = Dict[str, Any]
PolicyID = str


class Policy:
    def __init__(self, model, agent_id: str, policy_id: str):
        self.model = model
        self.agent_id = agent_id
        self.policy_id = policy_id
        self._state = self.get_initial_state()

    def step(self, obs):
        self._state = self.get_next_state(self._state.get("action"), obs, self._state)
        a = self.sample_action(self._state)
        self._state["action"] = a
        return a

    def reset(self, *, seed=None):
        self._state = self.get_initial_state()

    def get_initial_state(self) -> PolicyState:
        return {"action": None}

    def get_next_state(self, action, obs, state) -> PolicyState:
        return {"action": action}

    def sample_action(self, state):
        raise NotImplementedError

    def get_pi(self, state):
        raise NotImplementedError

    def get_value(self, state):
        raise NotImplementedError

    def close(self):
        pass


class RandomPolicy(Policy):
    def __init__(self, model, agent_id, policy_id="Random-v0"):
        super().__init__(model, agent_id, policy_id)
        self._space = model.action_spaces[agent_id]

    def sample_action(self, state):
        return self._space.sample()
