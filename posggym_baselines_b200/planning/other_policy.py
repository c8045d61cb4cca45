"""Other-agent policy plugin interface -- the reference's ``OtherAgentPolicy`` ABC,
``RandomOtherAgentPolicy`` and ``OtherAgentMixturePolicy``
(/root/reference/posggym_baselines/planning/other_policy.py:10-229).  ``RandomOtherAgentPolicy``
is executed inside the kernels; every other class drives the planner through
planning/host_search.py."""
from __future__ import annotations

import abc
import random
from typing import Any, Dict, List, Optional

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

    def get_state_from_history(self, initial_state: PolicyState, history) -> PolicyState:
        """Unroll the policy over an agent's ``(action, obs)`` history (other_policy.py:95-122)."""
        state = initial_state
        for action, obs in history:
            state = self.get_next_state(action, obs, state)
        return state

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


def _full_support(probs: Dict[Any, float], n_actions: int) -> Dict[Any, float]:
    """A ``posggym.agents`` action distribution may omit zero-probability actions; the planner indexes all."""
    if len(probs) != n_actions:
        for a in range(n_actions):
            probs.setdefault(a, 0.0)
    return probs


class OtherAgentMixturePolicy(OtherAgentPolicy):
    """The other agent follows ONE of a set of ``posggym.agents``-style policies for the whole episode
    (other_policy.py:157-229): the active one is drawn uniformly in ``sample_initial_state`` and its id
    rides in the policy state (``{"policy_id", "policy_state"}``), which is what the POTMMCP meta-policy keys
    on (planning/potmmcp.py).  Two agents only, as in the reference."""

    def __init__(self, model, agent_id: str, policies: Dict[str, Any]):
        super().__init__(model, agent_id)
        assert len(model.possible_agents) == 2, "Currently only supports 2 agents"
        self.policies = policies
        self.action_space = list(range(model.action_spaces[agent_id].n))

    def sample_initial_state(self) -> PolicyState:
        pid = random.choice(list(self.policies))
        return {"policy_id": pid, "policy_state": self.policies[pid].get_initial_state()}

    def get_next_state(self, action: Optional[Any], obs, state: PolicyState) -> PolicyState:
        pid = state["policy_id"]
        return {"policy_id": pid, "policy_state": self.policies[pid].get_next_state(action, obs, state["policy_state"])}

    def sample_action(self, state: PolicyState):
        return self.policies[state["policy_id"]].sample_action(state["policy_state"])

    def get_pi(self, state: PolicyState) -> Dict[Any, float]:
        return _full_support(self.policies[state["policy_id"]].get_pi(state["policy_state"]).probs, len(self.action_space))

    def close(self):
        for policy in self.policies.values():
            policy.close()

    @staticmethod
    def load_posggym_agents_policy(model, agent_id: str, policy_ids: List[str]) -> "OtherAgentMixturePolicy":
        import posggym.agents as pga   # only where the real posggym.agents registry is installed

        return OtherAgentMixturePolicy(model, agent_id, {pid: pga.make(pid, model, agent_id) for pid in policy_ids})
