"""Planner facade with the reference's names (posggym_baselines/planning/__init__.py)."""
from posggym_baselines_b200.planning.config import KnownBounds, MCTSConfig  # noqa: F401
from posggym_baselines_b200.planning.intmcp import INTMCP  # noqa: F401
from posggym_baselines_b200.planning.ipomcp import IPOMCP  # noqa: F401
from posggym_baselines_b200.planning.mcts import MCTS, POMCP  # noqa: F401
from posggym_baselines_b200.planning.other_policy import (  # noqa: F401
    OtherAgentMixturePolicy,
    OtherAgentPolicy,
    RandomOtherAgentPolicy,
)
from posggym_baselines_b200.planning.potmmcp import POTMMCP, POTMMCPMetaPolicy  # noqa: F401
from posggym_baselines_b200.planning.search_policy import (  # noqa: F401
    PPOLSTMSearchPolicy,
    RandomSearchPolicy,
    SearchPolicy,
    SearchPolicyWrapper,
    load_posggym_agents_search_policy,
)
from posggym_baselines_b200.planning.utils import PlanningStatTracker  # noqa: F401
