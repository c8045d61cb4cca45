"""``PlanningStatTracker`` with the reference's keys and aggregation
(/root/reference/posggym_baselines/planning/utils.py:45-144): per-episode means, max for
``mem_usage``; ``num_sims`` per decision is BASELINE.json's planner metric.  ``BeliefStatTracker``
(utils.py:147-339): how much of the root belief agrees with what really happened."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

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


def _particle_fields(p):
    """(state, history, policy_state) of a root particle: the plugin path keeps ``HPS`` tuples
    (planning/host_search.py), the device search packed states only (``state_belief_only`` semantics)."""
    if hasattr(p, "policy_state") and hasattr(p, "history"):
        return p.state, p.history, p.policy_state
    return p, None, None


class BeliefStatTracker:
    """Accuracy of the planner's root belief against the environment, per episode (utils.py:147-339):

    - ``state``: share of particles whose state is the true state before the most recent action;
    - ``history``: share whose other-agent histories equal the tracked ones;
    - ``policy``: share whose other-agent ``policy_id`` is the id of the policy really played;
    - ``action``: mean probability the particles' other-agent policies give to the actions just taken.

    ``env`` is an agent-folding env wrapper (``envs/agent_env.py``: ``policies``, ``last_actions``, ``last_obs``,
    ``state``).  Call ``reset`` / ``reset_episode`` after ``env.reset()`` and ``step(env)`` after ``planner.step``
    and ``env.step``.  Values and keys are the reference's -- including that the tracked other-agent histories stay
    at the episode's initial observation (the reference discards the result of the immutable ``extend``,
    utils.py:245-246), so ``history`` can only be non-zero on an episode's first step.
    """

    TRACKABLE_BELIEF_STATS = ["policy", "action", "state", "history"]

    def __init__(self, planner, env, track_per_step: bool = False,
                 stats_to_track: Sequence[str] = ("policy", "action", "state", "history")):
        stats_to_track = list(stats_to_track)
        assert all(k in self.TRACKABLE_BELIEF_STATS for k in stats_to_track)
        if "policy" in stats_to_track or "action" in stats_to_track:
            from posggym_baselines_b200.planning.other_policy import OtherAgentPolicy

            assert all(isinstance(pi, OtherAgentPolicy) for pi in planner.other_agent_policies.values())
        self.step_limit: Optional[int] = None
        if track_per_step:
            assert env.spec is not None and env.spec.max_episode_steps is not None
            self.step_limit = env.spec.max_episode_steps
        self.planner, self.env = planner, env
        self.track_per_step, self.stats_to_track = track_per_step, stats_to_track
        self.other_agent_ids = [i for i in planner.model.possible_agents if i != planner.agent_id]
        self._steps, self._state, self._num_episodes = 0, None, 0
        self._histories = {i: () for i in self.other_agent_ids}
        self._stats: Dict[str, List[float]] = {k: [] for k in stats_to_track}
        self.reset()

    @classmethod
    def get_stat_keys(cls, stats_to_track: Sequence[str], track_per_step: bool, step_limit: Optional[int]) -> List[str]:
        """``belief_<k>_acc`` (+ ``belief_<k>_acc_<t>`` per step), the columns of ``episode_results.csv``."""
        assert all(k in cls.TRACKABLE_BELIEF_STATS for k in stats_to_track)
        keys = []
        for k in stats_to_track:
            keys.append(f"belief_{k}_acc")
            if track_per_step:
                assert step_limit is not None
                keys.extend(f"belief_{k}_acc_{t}" for t in range(step_limit))
        return keys

    def step(self, env):
        self._steps += 1
        measure = {"policy": lambda: self.get_other_policy_accuracy(env.policies),
                   "action": lambda: self.get_other_action_accuracy(env.last_actions),    # the actions just performed
                   "state": lambda: self.get_state_accuracy(self._state),               # state before that action
                   "history": lambda: self.get_other_history_accuracy(self._histories)}
        empty = self.planner.root.belief.size() == 0
        for k in self.stats_to_track:
            self._stats[k].append(0.0 if empty else measure[k]())
        self._state = env.state

    def reset(self):
        self.reset_episode()
        self._num_episodes = 0

    def reset_episode(self):
        assert all(a is None for a in self.env.last_actions.values()), \
            "Must call reset environment before resetting BeliefStatTracker."
        self._histories = {i: ((None, o),) for i, o in self.env.last_obs.items()}
        self._state = self.env.state
        if self._steps == 0:
            return
        self._num_episodes += 1
        self._steps = 0
        self._stats = {k: [] for k in self.stats_to_track}

    def get_episode(self) -> Dict[str, float]:
        out = {}
        for k, v in self._stats.items():
            out[f"belief_{k}_acc"] = np.nan if len(v) == 0 or np.isnan(np.sum(v)) else np.nanmean(v, axis=0)
            if self.track_per_step:
                for t in range(self.step_limit):
                    out[f"belief_{k}_acc_{t}"] = v[t] if t < len(v) else np.nan
        return out

    # ------------------------------------------------------------------ the four measures
    def _particles(self):
        return [_particle_fields(p) for p in self.planner.root.belief.particles]

    def get_state_accuracy(self, state) -> float:
        return sum(int(s == state) for s, _, _ in self._particles()) / self.planner.root.belief.size()

    def get_other_history_accuracy(self, histories: Dict[str, tuple]) -> float:
        """``histories[i]`` = ``((None, o0), (a1, o1), ...)``; a particle's joint history is a tuple of
        (joint action, joint observation) pairs."""
        n = 0
        for _, hist, _ in self._particles():
            if hist is None:
                continue
            n += all(tuple((ja[i], jo[i]) for ja, jo in hist) == tuple(h) for i, h in histories.items())
        return n / self.planner.root.belief.size()

    def get_other_policy_accuracy(self, policies: Dict[str, object]) -> float:
        n = 0
        for _, _, pstate in self._particles():
            if pstate is None:
                continue
            n += all(st.get("policy_id", None) == policies[i].policy_id
                     for i, st in pstate.items() if i != self.planner.agent_id)
        return n / self.planner.root.belief.size()

    def get_other_action_accuracy(self, actions: Dict[str, int]) -> float:
        total = 0
        for _, _, pstate in self._particles():
            if pstate is None:
                continue
            prob = 1.0
            for i, a in actions.items():
                prob *= self.planner.other_agent_policies[i].get_pi(pstate[i]).get(a, 0.0)
            total += prob
        return total / self.planner.root.belief.size()
