"""Search / rollout policy plugin interface -- same names, arguments and error
behaviour as the reference's ``SearchPolicy`` ABC and ``RandomSearchPolicy``
(/root/reference/posggym_baselines/planning/search_policy.py:16-185).  The B200
planners recognise ``RandomSearchPolicy`` by type and execute it inside the kernels
(uniform actions from Philox); the scalar methods remain for code that calls them."""
from __future__ import annotations

import abc
from typing import Any, Dict, Optional

PolicyState = Dict[str, Any]


class SearchPolicy(abc.ABC):
    """A policy used to guide the search (and rollouts) of a planner."""

    def __init__(self, model, agent_id: str, policy_id: str):
        self.model = model
        self.agent_id = agent_id
        self.policy_id = policy_id

    @abc.abstractmethod
    def get_initial_state(self) -> PolicyState:
        """Get the initial policy state."""

    @abc.abstractmethod
    def get_next_state(self, action, obs, state: PolicyState) -> PolicyState:
        """Get the next policy state given the last action and observation."""

    @abc.abstractmethod
    def sample_action(self, state: PolicyState):
        """Sample an action given the policy state."""

    @abc.abstractmethod
    def get_pi(self, state: PolicyState) -> Dict[Any, float]:
        """Get the policy's distribution over actions for the policy state."""

    @abc.abstractmethod
    def get_value(self, state: PolicyState) -> float:
        """Value estimate; raise NotImplementedError when unsupported (the planner
        then falls back to a rollout, planning/mcts.py:392-399)."""

    def get_state_from_history(self, history) -> PolicyState:
        state = self.get_initial_state()
        for action, obs in history:
            state = self.get_next_state(action, obs, state)
        return state

    def close(self):
        """Close policy and perform any necessary cleanup."""


class RandomSearchPolicy(SearchPolicy):
    """Uniform random search policy."""

    def __init__(self, model, agent_id: str):
        super().__init__(model, agent_id, "RandomSearchPolicy")
        self._action_space = model.action_spaces[agent_id]

    def get_initial_state(self) -> PolicyState:
        return {}

    def get_next_state(self, action: Optional[Any], obs, state: PolicyState) -> PolicyState:
        return {}

    def sample_action(self, state: PolicyState):
        return self._action_space.sample()

    def get_pi(self, state: PolicyState) -> Dict[Any, float]:
        return {a: 1.0 / self._action_space.n for a in range(self._action_space.n)}

    def get_value(self, state: PolicyState) -> float:
        raise NotImplementedError("RandomSearchPolicy does not support value estimates.")
