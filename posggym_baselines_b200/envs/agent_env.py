"""Folding the other agents into the environment: the planner's experiment loop steps the env with its own
action only (baseline_exps/exp_utils.py:474-477,511-514).  The reference takes ``AgentEnvWrapper`` from
``posggym.agents.wrappers`` (third-party, absent from /root/reference -- restated here from what the reference
reads of it: ``policies``, ``last_actions``, ``last_obs``, ``state``, ``model``, ``spec``; planning/utils.py:
224-262) and ``UniformOtherAgentFn`` from posggym_baselines/utils/agent_env_wrapper.py:8-30."""
from __future__ import annotations

from typing import Callable, Dict, List, Optional


class UniformOtherAgentFn:
    """Draws each folded agent's policy uniformly (``model.rng.choice``) from its list of policy ids at every
    reset; policy objects are made once and cached.  ``make_policy(policy_id, model, agent_id)`` builds one
    (``posggym.agents.make`` where that registry is installed)."""

    def __init__(self, agent_policy_ids: Dict[str, List[str]], make_policy: Optional[Callable] = None):
        self.agent_policy_ids = agent_policy_ids
        self.policies: Dict[str, Dict[str, object]] = {i: {} for i in agent_policy_ids}
        self._make = make_policy

    def __call__(self, model) -> Dict[str, object]:
        make = self._make
        if make is None:
            import posggym.agents as pga   # only where the real registry is installed

            make = pga.make
        chosen = {}
        for agent_id, ids in self.agent_policy_ids.items():
            pid = model.rng.choice(ids)
            if pid not in self.policies[agent_id]:
                self.policies[agent_id][pid] = make(pid, model, agent_id)
            chosen[agent_id] = self.policies[agent_id][pid]
        return chosen


class AgentEnvWrapper:
    """``env`` with the agents ``agent_fn(model)`` returns played by those policies: ``step`` takes the remaining
    agents' actions.  ``last_obs`` / ``last_actions`` hold the folded agents' most recent observations and actions
    (actions ``None`` right after a reset)."""

    def __init__(self, env, agent_fn: Callable):
        self.env, self.agent_fn = env, agent_fn
        self.policies: Dict[str, object] = {}
        self.last_obs: Dict[str, object] = {}
        self.last_actions: Dict[str, Optional[int]] = {}
        self._pstate: Dict[str, object] = {}

    def __getattr__(self, name):          # model, possible_agents, action_spaces, state, ...
        return getattr(self.env, name)

    @property
    def spec(self):
        return getattr(self.env, "spec", None) or getattr(self.env.model, "spec", None)

    def reset(self, seed: Optional[int] = None, options=None):
        obs, infos = self.env.reset(seed=seed)
        self.policies = self.agent_fn(self.env.model)
        self._pstate = {i: pi.get_initial_state() for i, pi in self.policies.items()}
        self.last_obs = {i: obs[i] for i in self.policies}
        self.last_actions = {i: None for i in self.policies}
        return obs, infos

    def step(self, actions: Dict[str, int]):
        joint = dict(actions)
        for i, pi in self.policies.items():
            self._pstate[i] = pi.get_next_state(self.last_actions[i], self.last_obs[i], self._pstate[i])
            joint[i] = pi.sample_action(self._pstate[i])
        out = self.env.step(joint)
        self.last_actions = {i: joint[i] for i in self.policies}
        self.last_obs = {i: out[0][i] for i in self.policies}
        return out

    def close(self):
        for pi in self.policies.values():
            pi.close()
        self.env.close()
