"""Other-agent policy plugin interface -- the reference's ``OtherAgentPolicy`` ABC and
``RandomOtherAgentPolicy`` (/root/reference/posggym_baselines/planning/other_policy.py:
10-154).  ``RandomOtherAgentPolicy`` is executed inside the kernels."""
from __future__ import annotations

import abc
from typing import Any, Dict, Optional

PolicyState = Dict[str, Any]


class OtherAgentPolicy(abc.ABC):
    """A policy for the other agent(s) in the environment, as believed by the planner."""

    def __init__(self, model, agent_id: str):
        self.model = model
        self.agent_id = agent_id

    @abc.abstractmethod
    def sample_initial_state(self) -> PolicyState:
        """Sample an initial policy state."""

    @abc.abstractmethod
    def get_next_state(self, action, obs, state: PolicyState) -> PolicyState:
        """Get the next policy state."""

    @abc.abstractmethod
    def sample_action(self, state: PolicyState):
        """Sample an action."""

    @abc.abstractmethod
    def get_pi(self, state: PolicyState) -> Dict[Any, float]:
        """Get the distribution over actions."""

    def close(self):
        """Close policy and perform any necessary cleanup."""


class RandomOtherAgentPolicy(OtherAgentPolicy):
    """Uniform random policy."""

    def __init__(self, model, agent_id: str):
        super().__init__(model, agent_id)
        self._action_space = model.action_spaces[agent_id]

    def sample_initial_state(self) -> PolicyState:
        return {}

    def get_next_state(self, action: Optional[Any], obs, state: PolicyState) -> PolicyState:
        return {}

    def sample_action(self, state: PolicyState):
        return self._action_space.sample()

    def get_pi(self, state: PolicyState) -> Dict[Any, float]:
        return {a: 1.0 / self._action_space.n for a in range(self._action_space.n)}
