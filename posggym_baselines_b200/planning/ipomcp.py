"""I-POMCP with known other-agent policies (reference: planning/ipomcp.py:10-38): ``MCTS`` under another
name -- history / policy-state particles, rejection-sampling belief update.  Which engine runs follows from
the plugins exactly as for ``MCTS`` (planning/mcts.py): the Random pair inside the kernels, anything else
through planning/host_search.py with the model arithmetic on the device."""
from __future__ import annotations

from typing import Dict

from posggym_baselines_b200.planning.config import MCTSConfig
from posggym_baselines_b200.planning.mcts import MCTS
from posggym_baselines_b200.planning.other_policy import OtherAgentPolicy
from posggym_baselines_b200.planning.search_policy import SearchPolicy


class IPOMCP(MCTS):
    """Interactive POMCP: MCTS with a belief over (state, joint history, other agents' policy states)."""

    def __init__(self, model, agent_id: str, config: MCTSConfig, other_agent_policies: Dict[str, OtherAgentPolicy],
                 search_policy: SearchPolicy, **b200_options):
        super().__init__(model, agent_id, config, other_agent_policies, search_policy, **b200_options)
