"""POTMMCP: type-based planning with a meta-policy over search policies
(reference: /root/reference/posggym_baselines/planning/potmmcp.py:18-301 the planner, 304-552 the meta-policy).

The other agent is believed to follow one of a set of policies (``OtherAgentMixturePolicy``; its id rides in
every particle's policy state).  Per simulation the meta-policy maps the particle's other-agent policy id to a
distribution over the planning agent's own policies, draws one, and that policy supplies the prior of the node
the simulation adds and its leaf evaluation; root / update-time nodes get the meta-policy's expected action
distribution under a uniform prior over the other agent's policies.

The policies are stateful Python objects, so the descent runs on the plugin path
(planning/host_search.py: initial belief and reinvigoration rounds batched on the device, model steps through
the facade); this file holds only what differs from ``MCTS``.
"""
from __future__ import annotations

import math
import random
from typing import Any, Dict, Optional

from posggym_baselines_b200.planning.config import MCTSConfig
from posggym_baselines_b200.planning.host_search import HostSearch
from posggym_baselines_b200.planning.mcts import MCTS
from posggym_baselines_b200.planning.other_policy import OtherAgentPolicy
from posggym_baselines_b200.planning.search_policy import PolicyState, SearchPolicy, SearchPolicyWrapper


class _PotmmcpEngine(HostSearch):
    """``HostSearch`` with the four POTMMCP differences (potmmcp.py:164-301)."""

    def _sim_policy(self, hps):
        # potmmcp.py:182-186: one policy per simulation, drawn for the particle's other-agent type
        return self.search_policy.sample_policy(hps.policy_state)

    def _node_prior(self, sp_state, sim_policy):
        if sim_policy is None:
            # update path (potmmcp.py:44-56, 121-139): expectation under a uniform prior over the types
            return self.search_policy.get_expected_action_probs(None, sp_state)
        return sim_policy.get_pi(sp_state[sim_policy.policy_id])          # potmmcp.py:266-276

    def _leaf_state(self, node, sim_policy):
        return node.search_policy_state[sim_policy.policy_id]             # potmmcp.py:211-216

    def _revisit(self, node, child, sim_policy):
        # potmmcp.py:244-252: the reference folds the sampled policy's distribution into the node the action
        # was selected AT (its visits as the averaging count), not into the child
        child.visits += 1
        pi = sim_policy.get_pi(node.search_policy_state[sim_policy.policy_id])
        for a, pr in pi.items():
            node.action_probs[a] += (pr - node.action_probs[a]) / node.visits


class POTMMCP(MCTS):
    """Partially Observable Type-Based Multi-Agent Monte-Carlo Planning."""

    def __init__(self, model, agent_id: str, config: MCTSConfig, other_agent_policies: Dict[str, OtherAgentPolicy],
                 search_policy: "POTMMCPMetaPolicy", **b200_options):
        if config.state_belief_only:
            raise ValueError("POTMMCP reads the other agent's policy id from the particles: state_belief_only=False")
        super().__init__(model, agent_id, config, other_agent_policies, search_policy, **b200_options)
        assert self.num_agents == 2, "Currently only supports 2 agents"
        assert self._host is not None

    def _host_engine(self):
        return _PotmmcpEngine


class POTMMCPMetaPolicy(SearchPolicy):
    """Search policy of POTMMCP (potmmcp.py:304-552): ``meta_policy[other agent's policy id]`` is a distribution
    over ``policies`` (the planning agent's candidate policies, ``posggym.agents``-style objects).  Its state is
    the dict of every candidate's state.  It is used through ``sample_policy`` (per simulation) and
    ``get_expected_action_probs`` (root prior); the scalar ``sample_action`` / ``get_pi`` / ``get_value`` refuse,
    as in the reference."""

    def __init__(self, model, agent_id: str, policies: Dict[str, Any], meta_policy: Dict[str, Dict[str, float]]):
        super().__init__(model, agent_id, "POTMMCPMetaPolicy")
        assert len(model.possible_agents) == 2, "Currently only supports 2 agents"
        assert len(meta_policy) > 0
        for dist in meta_policy.values():
            assert all(pid in policies for pid in dist)
            assert abs(sum(dist.values()) - 1) < 1e-6
        self.other_agent_id = next(i for i in model.possible_agents if i != agent_id)
        self.policies = policies
        self.meta_policy = meta_policy
        self.action_space = list(range(model.action_spaces[agent_id].n))

    def get_initial_state(self) -> PolicyState:
        return {pid: pi.get_initial_state() for pid, pi in self.policies.items()}

    def get_next_state(self, action: Optional[Any], obs, state: PolicyState) -> PolicyState:
        return {pid: self.policies[pid].get_next_state(action, obs, st) for pid, st in state.items()}

    def sample_action(self, state: PolicyState):
        raise NotImplementedError("POTMMCPMetaPolicy does not sample actions: sample_policy() first, then that policy")

    def get_pi(self, state: PolicyState) -> Dict[Any, float]:
        raise NotImplementedError("POTMMCPMetaPolicy has no action distribution: sample_policy() first, then that policy")

    def get_value(self, state: PolicyState) -> float:
        raise NotImplementedError("POTMMCPMetaPolicy has no value estimate: sample_policy() first, then that policy")

    def sample_policy(self, other_agent_policy_state: Dict[str, PolicyState]) -> SearchPolicy:
        """The search policy of one simulation given the other agent's policy state (one ``random.choices`` draw)."""
        dist = self.meta_policy[other_agent_policy_state[self.other_agent_id]["policy_id"]]
        pid = random.choices(list(dist), weights=list(dist.values()), k=1)[0]
        return SearchPolicyWrapper(self.policies[pid])

    def get_expected_action_probs(self, other_agent_policy_dist: Optional[Dict[str, float]],
                                  policy_state: PolicyState) -> Dict[Any, float]:
        """E over the other agent's policy (uniform when None) of the meta-policy's mixture of action distributions."""
        if other_agent_policy_dist is None:
            other_agent_policy_dist = dict.fromkeys(self.meta_policy, 1.0 / len(self.meta_policy))
        weight = dict.fromkeys(self.policies, 0.0)
        for other_id, p_other in other_agent_policy_dist.items():
            for pid, p_meta in self.meta_policy[other_id].items():
                weight[pid] += p_other * p_meta
        z = sum(weight.values())
        out = dict.fromkeys(self.action_space, 0.0)
        for pid, w in weight.items():
            w /= z
            for a, p in self.policies[pid].get_pi(policy_state[pid]).probs.items():
                out[a] += w * p
        z = sum(out.values())
        return {a: p / z for a, p in out.items()}

    def close(self):
        for pi in self.policies.values():
            pi.close()

    # ------------------------------------------------------------------ meta-policy construction
    @staticmethod
    def load_posggym_agents_meta_policy(model, agent_id: str, meta_policy: Dict[str, Dict[str, float]]) -> "POTMMCPMetaPolicy":
        import posggym.agents as pga   # only where the real posggym.agents registry is installed

        ids = {pid for dist in meta_policy.values() for pid in dist}
        return POTMMCPMetaPolicy(model, agent_id, {pid: pga.make(pid, model, agent_id) for pid in ids}, meta_policy)

    @staticmethod
    def _other_ids(pairwise_returns: Dict[str, Dict[str, float]]):
        return {oid for row in pairwise_returns.values() for oid in row}

    @staticmethod
    def get_uniform_meta_policy(pairwise_returns: Dict[str, Dict[str, float]]) -> Dict[str, Dict[str, float]]:
        """``pairwise_returns[ego policy][other policy]`` -> every ego policy equally likely against every type."""
        n = len(pairwise_returns)
        return {oid: {pid: 1.0 / n for pid in pairwise_returns} for oid in POTMMCPMetaPolicy._other_ids(pairwise_returns)}

    @staticmethod
    def get_greedy_meta_policy(pairwise_returns: Dict[str, Dict[str, float]]) -> Dict[str, Dict[str, float]]:
        """All mass, split evenly, on the ego policies with the best return against each type."""
        out = {}
        for oid in POTMMCPMetaPolicy._other_ids(pairwise_returns):
            best = max(row[oid] for row in pairwise_returns.values())
            winners = [pid for pid, row in pairwise_returns.items() if row[oid] == best]
            out[oid] = {pid: (1.0 / len(winners) if pid in winners else 0.0) for pid in pairwise_returns}
        return out

    @staticmethod
    def get_softmax_meta_policy(pairwise_returns: Dict[str, Dict[str, float]], temperature: float) -> Dict[str, Dict[str, float]]:
        """Softmax of the returns against each type at ``temperature``."""
        out = {}
        for oid in POTMMCPMetaPolicy._other_ids(pairwise_returns):
            e = [math.e ** (row[oid] / temperature) for row in pairwise_returns.values()]
            z = sum(e)
            out[oid] = {pid: v / z for pid, v in zip(pairwise_returns, e)}
        return out
