"""The result format on the far side of the planning path: ``run_planning_exp`` of the
reference's experiment harness (/root/reference/baseline_exps/exp_utils.py:468-552) -- the episode loop that
drives ``planner.step`` and appends one row per episode to ``episode_results.csv`` -- with the reference's columns
in the reference's order (exp_utils.py:325-342): ``num, len, return, discounted_return, time`` followed by
``PlanningStatTracker.STAT_KEYS`` (planning/utils.py:49-61) and, when asked for, belief statistics.  A file written
here loads into the reference's analysis notebooks next to the committed ``*_results.csv``.

Only the loop and the file format: experiment grids, process pools, argument files and logs are the reference's
control plane and stay there."""
from __future__ import annotations

import csv
import os
import time
from typing import Callable, Dict, List, Optional, Sequence

from posggym_baselines_b200.planning.utils import BeliefStatTracker, PlanningStatTracker

EPISODE_KEYS = ["num", "len", "return", "discounted_return", "time"]


def episode_results_heads(extra: Sequence[str] = ()) -> List[str]:
    """Column order of ``episode_results.csv`` (exp_utils.py:325-342)."""
    return EPISODE_KEYS + list(PlanningStatTracker.STAT_KEYS) + list(extra)


def run_planning_exp(env, planner, agent_id: str, other_agent_fn: Optional[Callable[[Dict], Dict[str, int]]],
                     num_episodes: int, results_file: Optional[str] = None, exp_time_limit: float = float("inf"),
                     extra_stats_fn: Optional[Callable[[], Dict[str, float]]] = None,
                     extra_keys: Sequence[str] = (), seed: Optional[int] = None,
                     belief_stats_to_track: Sequence[str] = (),
                     track_per_step_belief_stats: bool = False) -> List[Dict[str, float]]:
    """Runs up to ``num_episodes`` episodes of ``planner`` (any object with the planner facade: ``reset``, ``step``,
    ``stat_tracker``, ``config.discount``) in ``env``.  The other agents' actions come from ``other_agent_fn(obs)``,
    or -- ``other_agent_fn=None`` -- ``env`` is an ``envs.agent_env.AgentEnvWrapper`` that has them folded in, as in
    the reference (exp_utils.py:474-477).  With ``belief_stats_to_track`` (needs the wrapper) a ``BeliefStatTracker``
    adds its ``belief_*_acc`` columns after the planner's (exp_utils.py:334-342,479-484,523,531).  Appends one row
    per episode to ``results_file`` (header written first) and returns the rows."""
    belief_tracker, belief_keys = None, []
    if belief_stats_to_track:
        belief_tracker = BeliefStatTracker(planner, env, track_per_step=track_per_step_belief_stats,
                                           stats_to_track=list(belief_stats_to_track))
        belief_keys = BeliefStatTracker.get_stat_keys(list(belief_stats_to_track), track_per_step_belief_stats,
                                                      belief_tracker.step_limit)
    heads = episode_results_heads(belief_keys + list(extra_keys))
    if results_file is not None:
        os.makedirs(os.path.dirname(os.path.abspath(results_file)), exist_ok=True)
        with open(results_file, "w", newline="") as f:
            csv.DictWriter(f, fieldnames=heads).writeheader()
    planner.stat_tracker.track_overall = False      # per-episode logging only (exp_utils.py:486)
    rows, start, episode = [], time.time(), 0
    while episode < num_episodes and time.time() - start < exp_time_limit:
        obs, _ = env.reset(seed=None if seed is None else seed + episode)
        planner.reset()
        if belief_tracker is not None:
            belief_tracker.reset()
        row = {"num": episode, "len": 0, "return": 0, "discounted_return": 0, "time": 0}
        t0, done = time.time(), False
        while not done:
            action = planner.step(obs[agent_id])
            joint = dict(other_agent_fn(obs)) if other_agent_fn is not None else {}
            joint[agent_id] = action
            obs, rewards, terms, truncs, all_done, _ = env.step(joint)
            done = terms[agent_id] or truncs[agent_id] or all_done      # only the planning agent's end counts
            r = rewards[agent_id]
            row["return"] += r
            row["discounted_return"] += planner.config.discount ** row["len"] * r
            row["len"] += 1
            if belief_tracker is not None:
                belief_tracker.step(env)
        row["time"] = time.time() - t0
        row.update(planner.stat_tracker.get_episode())
        if belief_tracker is not None:
            row.update(belief_tracker.get_episode())
        if extra_stats_fn is not None:
            row.update(extra_stats_fn())
        if results_file is not None:
            with open(results_file, "a", newline="") as f:
                csv.DictWriter(f, fieldnames=heads).writerow({k: row.get(k) for k in heads})
        rows.append(row)
        episode += 1
    return rows
