"""posggym.model stand-in: the names the reference's evaluated annotations and
call sites need (mcts.py:97,159,175,333-342)."""
from __future__ import annotations

import abc
import enum
import random
from typing import Any, Dict, Generic, List, NamedTuple, Optional, TypeVar

StateType = TypeVar("StateType")
ObsType = TypeVar("ObsType")
ActType = TypeVar("ActType")
AgentID = str


class Outcome(enum.Enum):
    LOSS = -1
    DRAW = 0
    WIN = 1
    NA = None

    def __str__(self):
        return self.name


class JointTimestep(NamedTuple):
    state: Any
    observations: Dict[str, Any]
    rewards: Dict[str, float]
    terminations: Dict[str, bool]
    truncations: Dict[str, bool]
    all_done: bool
    infos: Dict[str, Dict]


class _Spec:
    def __init__(self, id: str, max_episode_steps: Optional[int]):
        self.id = id
        self.max_episode_steps = max_episode_steps


class POSGModel(abc.ABC, Generic[StateType, ObsType, ActType]):
    possible_agents: tuple
    action_spaces: Dict[str, Any]
    observation_spaces: Dict[str, Any]
    is_symmetric: bool = True
    spec: Optional[_Spec] = None
    _rng: Optional[random.Random] = None

    @property
    def rng(self) -> random.Random:
        if self._rng is None:
            self._rng = random.Random()
        return self._rng

    def seed(self, seed: Optional[int] = None):
        self._rng = random.Random(seed)
        for i, sp in enumerate(self.action_spaces.values()):
            sp.seed(None if seed is None else seed + 1 + i)

    def get_agents(self, state) -> List[str]:
        return list(self.possible_agents)

    @abc.abstractmethod
    def sample_initial_state(self):
        ...

    @abc.abstractmethod
    def sample_initial_obs(self, state):
        ...

    def sample_agent_initial_state(self, agent_id: str, obs):
        raise NotImplementedError

    @abc.abstractmethod
    def step(self, state, actions) -> JointTimestep:
        ...
