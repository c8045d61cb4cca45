"""``PlanningStatTracker`` with the reference's keys and aggregation
(/root/reference/posggym_baselines/planning/utils.py:45-144): per-episode means, max for
``mem_usage``; ``num_sims`` per decision is BASELINE.json's planner metric."""
from __future__ import annotations

from typing import Dict, List

import numpy as np

from posggym_baselines_b200.planning.config import KnownBounds  # noqa: F401


class PlanningStatTracker:
    """Tracks MCTS planning statistics."""

    STAT_KEYS = [
        "search_time", "update_time", "reinvigoration_time", "evaluation_time", "policy_calls",
        "inference_time", "search_depth", "num_sims", "mem_usage", "min_value", "max_value",
    ]
    MAX_STATS = {"mem_usage"}

    def __init__(self, planner, track_overall: bool = True):
        self.planner = planner
        self.track_overall = track_overall
        self._current_steps = 0
        self._current_stats: Dict[str, List[float]] = {k: [] for k in self.STAT_KEYS}
        self._num_episodes = 0
        self._all_steps: List[int] = []
        self._all_stats: Dict[str, List[float]] = {k: [] for k in self.STAT_KEYS}
        self.reset()

    def step(self):
        self._current_steps += 1
        for k in self.STAT_KEYS:
            self._current_stats[k].append(self.planner.step_statistics.get(k, np.nan))

    def reset(self):
        self.reset_episode()
        self._num_episodes = 0
        self._all_steps = []
        self._all_stats = {k: [] for k in self.STAT_KEYS}

    def reset_episode(self):
        if self._current_steps == 0:
            return
        self._num_episodes += 1
        if self.track_overall:
            self._all_steps.append(self._current_steps)
            for k in self.STAT_KEYS:
                agg = np.nanmax if k in self.MAX_STATS else np.nanmean
                self._all_stats[k].append(agg(self._current_stats[k]))
        self._current_steps = 0
        self._current_stats = {k: [] for k in self.STAT_KEYS}

    def get_episode(self) -> Dict[str, float]:
        stats = {}
        for key, vals in self._current_stats.items():
            if len(vals) == 0 or np.isnan(np.sum(vals)):
                stats[key] = np.nan
            elif key in self.MAX_STATS:
                stats[key] = np.nanmax(vals, axis=0)
            else:
                stats[key] = np.nanmean(vals, axis=0)
        return stats

    def get(self) -> Dict[str, float]:
        stats = {}
        for key, values in self._all_stats.items():
            if len(values) == 0 or np.isnan(np.sum(values)):
                mean_val = std_val = np.nan
            elif key in self.MAX_STATS:
                mean_val, std_val = np.nanmax(values, axis=0), np.nanstd(values, axis=0)
            else:
                mean_val, std_val = np.nanmean(values, axis=0), np.nanstd(values, axis=0)
            stats[f"{key}_mean"] = mean_val
            stats[f"{key}_std"] = std_val
        return stats
