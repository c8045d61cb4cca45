"""Planner facade with the reference's names (posggym_baselines/planning/__init__.py)."""
from posggym_baselines_b200.planning.config import KnownBounds, MCTSConfig  # noqa: F401
from posggym_baselines_b200.planning.intmcp import INTMCP  # noqa: F401
from posggym_baselines_b200.planning.mcts import MCTS, POMCP  # noqa: F401
from posggym_baselines_b200.planning.other_policy import OtherAgentPolicy, RandomOtherAgentPolicy  # noqa: F401
from posggym_baselines_b200.planning.search_policy import PPOLSTMSearchPolicy, RandomSearchPolicy, SearchPolicy  # noqa: F401
from posggym_baselines_b200.planning.utils import PlanningStatTracker  # noqa: F401
