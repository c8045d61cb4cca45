"""posggym.utils.history stand-in (AgentHistory / JointHistory), API as consumed at
intmcp.py:101,206-210,220,283,293,314,322,354,800 and mcts.py:204-206,348."""
from __future__ import annotations

from typing import Dict, Optional, Tuple


class AgentHistory:
    """Immutable action-observation history of one agent: ((a0,o0),(a1,o1),...)."""

    def __init__(self, history: Tuple[tuple, ...]):
        self.history = tuple(history)
        self.t = len(self.history) - 1

    def extend(self, action, obs) -> "AgentHistory":
        return AgentHistory((*self.history, (action, obs)))

    def get_sub_history(self, horizon: int) -> "AgentHistory":
        if horizon >= len(self.history):
            return self
        return AgentHistory(self.history[:horizon])

    def get_last_step(self):
        return self.history[-1]

    @classmethod
    def get_init_history(cls, obs=None) -> "AgentHistory":
        if obs is None:
            return cls(())
        return cls(((None, obs),))

    def __hash__(self):
        return hash(self.history)

    def __eq__(self, other):
        return isinstance(other, AgentHistory) and self.history == other.history

    def __len__(self):
        return len(self.history)

    def __getitem__(self, i):
        return self.history[i]

    def __iter__(self):
        return iter(self.history)

    def __str__(self):
        return "<" + ",".join(f"({a},{o})" for a, o in self.history) + ">"

    __repr__ = __str__


class JointHistory:
    """Joint history: one AgentHistory per agent."""

    def __init__(self, agent_histories: Dict[str, AgentHistory]):
        self.agent_histories = agent_histories
        self.agent_ids = tuple(agent_histories)

    @classmethod
    def get_init_history(cls, agent_ids, obs: Optional[Dict[str, object]] = None) -> "JointHistory":
        if obs is None:
            return cls({i: AgentHistory.get_init_history() for i in agent_ids})
        return cls({i: AgentHistory.get_init_history(obs[i]) for i in agent_ids})

    def get_agent_history(self, agent_id: str) -> AgentHistory:
        return self.agent_histories[agent_id]

    def extend(self, action: Dict[str, object], obs: Dict[str, object]) -> "JointHistory":
        return JointHistory(
            {i: h.extend(action[i], obs[i]) for i, h in self.agent_histories.items()}
        )

    def __hash__(self):
        return hash(tuple(self.agent_histories.items()))

    def __eq__(self, other):
        return isinstance(other, JointHistory) and self.agent_histories == other.agent_histories
