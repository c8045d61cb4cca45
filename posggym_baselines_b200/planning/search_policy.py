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


class SearchPolicyWrapper(SearchPolicy):
    """A ``posggym.agents`` ``Policy`` object as a ``SearchPolicy`` (search_policy.py:188-232): same calls
    forwarded, ``get_pi`` unwrapped from the action distribution and padded to every action."""

    def __init__(self, policy):
        super().__init__(policy.model, policy.agent_id, policy.policy_id)
        self.policy = policy
        self.action_space = list(range(policy.model.action_spaces[policy.agent_id].n))

    def get_initial_state(self) -> PolicyState:
        return self.policy.get_initial_state()

    def get_next_state(self, action: Optional[Any], obs, state: PolicyState) -> PolicyState:
        return self.policy.get_next_state(action, obs, state)

    def sample_action(self, state: PolicyState):
        return self.policy.sample_action(state)

    def get_pi(self, state: PolicyState) -> Dict[Any, float]:
        probs = self.policy.get_pi(state).probs
        if len(probs) != len(self.action_space):
            for a in self.action_space:
                probs.setdefault(a, 0.0)
        return probs

    def get_value(self, state: PolicyState) -> float:
        return self.policy.get_value(state)

    def close(self):
        self.policy.close()


def load_posggym_agents_search_policy(model, agent_id: str, policy_id: str) -> SearchPolicyWrapper:
    import posggym.agents as pga   # only where the real posggym.agents registry is installed

    return SearchPolicyWrapper(pga.make(policy_id, model, agent_id))


def _flat_index_layout(space):
    """(offset, start) of every Discrete leaf of a (nested) Tuple space in gymnasium's ``flatten`` order, and the
    flattened width -- the one-hot layout ``processors.FlattenProcessor`` produces for the reference's policies."""
    leaves, width = [], 0

    def walk(sp):
        nonlocal width
        if hasattr(sp, "spaces"):
            for s in sp.spaces:
                walk(s)
        else:
            leaves.append((width, int(getattr(sp, "start", 0))))
            width += int(sp.n)

    walk(space)
    return leaves, width


def _flat_values(obs):
    if isinstance(obs, (tuple, list)):
        for o in obs:
            yield from _flat_values(o)
    else:
        yield int(obs)


class PPOLSTMSearchPolicy(SearchPolicy):
    """The reference's ``PPOLSTMSearchPolicy`` (planning/search_policy.py:235-288; built by
    baseline_exps/run_combined_exps.py:48-91 from a BR checkpoint): a trained ``PPOLSTMModel`` as search policy
    -- action prior for PUCB, value estimate for truncated search, LSTM state carried along the tree.

    Here the network runs on the GPU: trunk layer 0 as a sum of the weight columns of the observation's one-hot
    inputs, everything else in the tensor-core policy kernel (``posg_policy_step``, ppo/policy.py) -- also for
    the one-row calls of the planner's host path (``get_next_state``), and for whole batches of tree nodes
    through ``get_next_state_batch``; either way ONE library call with host rows (``posg_policy_eval_host``).
    ``state_dict`` uses the reference's parameter names (``checkpoint["model"]``); a policy state's LSTM pair is
    held as CPU tensors, as in the reference."""

    def __init__(self, model, agent_id: str, policy_id: str, state_dict, device=None):
        import torch

        from posggym_baselines_b200.ppo.policy import FusedPolicy

        super().__init__(model, agent_id, policy_id)
        self.action_space = list(range(model.action_spaces[agent_id].n))
        self._torch = torch
        dev = device if device is not None else getattr(getattr(model, "batched", None), "device", None)
        self.policy = FusedPolicy(state_dict, device=dev)
        self.device = self.policy.device
        self._leaves, width = _flat_index_layout(model.observation_spaces[agent_id])
        if width != self.policy.w0_t.shape[0]:
            raise ValueError(f"the policy expects {self.policy.w0_t.shape[0]} inputs, the observation space flattens to {width}")

    def get_initial_state(self) -> PolicyState:
        t = self._torch
        z = lambda: t.zeros((1, self.policy.hidden), dtype=t.float32)  # noqa: E731
        return {"lstm_state": (z(), z()), "pi": None, "value": 0.0}

    def _index_rows(self, obs_list):
        import numpy as np

        return np.array([[off + v - start for (off, start), v in zip(self._leaves, _flat_values(o))] for o in obs_list],
                        dtype=np.int32)

    def get_next_state_batch(self, obs_list, states):
        """``get_next_state`` for many (observation, policy state) pairs in one evaluation of the network
        (``posg_policy_eval_host``: host rows in, one launch of trunk layer 0 and one of the tensor-core policy
        kernel, host rows out).  The LSTM state of a policy state is a pair of CPU tensors [1, 64], as in the
        reference, whose search policy runs its network on the CPU."""
        import numpy as np

        t = self._torch
        n = len(obs_list)
        h = np.concatenate([s["lstm_state"][0].numpy() for s in states], 0)
        c = np.concatenate([s["lstm_state"][1].numpy() for s in states], 0)
        h2, c2, _, probs, values = self.policy.eval_host(self._index_rows(obs_list), h, c)   # probs: search_policy.py:270
        h2, c2 = t.from_numpy(h2), t.from_numpy(c2)
        return [{"lstm_state": (h2[k:k + 1], c2[k:k + 1]),
                 "pi": {a: float(probs[k, a]) for a in self.action_space}, "value": float(values[k])} for k in range(n)]

    def get_next_state(self, action, obs, state: PolicyState) -> PolicyState:
        return self.get_next_state_batch([obs], [state])[0]

    def sample_action(self, state: PolicyState):
        import random

        pi = state["pi"]
        return random.choices(list(pi.keys()), weights=list(pi.values()), k=1)[0]

    def get_pi(self, state: PolicyState) -> Dict[Any, float]:
        return state["pi"]

    def get_value(self, state: PolicyState) -> float:
        return state["value"]
