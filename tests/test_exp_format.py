"""episode_results.csv as the reference's harness writes it (baseline_exps/exp_utils.py:325-342,468-552)."""
import csv
import os

import numpy as np
import pytest

REF_CSV = "/root/reference/baseline_exps/env_data/Driving-v1/combined_belief_results.csv"
EXPECTED = ["num", "len", "return", "discounted_return", "time", "search_time", "update_time", "reinvigoration_time",
            "evaluation_time", "policy_calls", "inference_time", "search_depth", "num_sims", "mem_usage", "min_value",
            "max_value"]


def test_columns_are_the_references():
    from posggym_baselines_b200.exp import episode_results_heads

    assert episode_results_heads() == EXPECTED
    if os.path.exists(REF_CSV):      # the committed results: experiment-argument columns, then exactly these
        with open(REF_CSV) as f:
            head = next(csv.reader(f))
        i = head.index("num")
        assert head[i:i + len(EXPECTED)] == EXPECTED


def test_episode_loop_writes_one_row_per_episode(tmp_path):
    from posggym_baselines_b200.exp import run_planning_exp
    from posggym_baselines_b200.planning.utils import PlanningStatTracker

    class Env:
        def reset(self, seed=None):
            self.t = 0
            return {"0": 0, "1": 0}, {}

        def step(self, a):
            self.t += 1
            done = self.t >= 4
            return {"0": self.t, "1": self.t}, {"0": 0.5, "1": 0.0}, {"0": done, "1": False}, {"0": False, "1": False}, done, {}

    class Planner:
        class config:
            discount = 0.9

        def __init__(self):
            self.step_statistics = {}
            self.stat_tracker = PlanningStatTracker(self)

        def reset(self):
            self.stat_tracker.reset_episode()

        def step(self, obs):
            self.step_statistics = {k: 1.0 for k in PlanningStatTracker.STAT_KEYS}
            self.step_statistics["num_sims"] = 10 + obs
            self.stat_tracker.step()
            return 1

    out = tmp_path / "episode_results.csv"
    rows = run_planning_exp(Env(), Planner(), "0", lambda obs: {"1": 0}, 3, str(out))
    assert len(rows) == 3
    with open(out) as f:
        got = list(csv.DictReader(f))
    assert list(got[0].keys()) == EXPECTED and len(got) == 3
    assert [int(r["num"]) for r in got] == [0, 1, 2] and all(int(r["len"]) == 4 for r in got)
    assert abs(float(got[0]["return"]) - 2.0) < 1e-12
    assert abs(float(got[0]["discounted_return"]) - 0.5 * (1 + 0.9 + 0.81 + 0.729)) < 1e-12
    assert abs(float(got[0]["num_sims"]) - np.mean([10, 11, 12, 13])) < 1e-12


@pytest.mark.gpu
def test_pomcp_episodes_to_csv(tmp_path):
    import random

    from posggym_baselines_b200.envs.env import make
    from posggym_baselines_b200.exp import run_planning_exp
    from posggym_baselines_b200.planning import POMCP, MCTSConfig, RandomSearchPolicy

    env = make("Driving-v1")
    cfg = MCTSConfig(discount=0.95, search_time_limit=0.1, c=1.414, truncated=False, action_selection="ucb", seed=0)
    planner = POMCP(env.model, "0", cfg, RandomSearchPolicy(env.model, "0"), max_sims=64, node_capacity=1 << 16,
                    log_capacity=1 << 19)
    rng = random.Random(0)
    out = tmp_path / "episode_results.csv"
    rows = run_planning_exp(env, planner, "0", lambda obs: {"1": rng.randrange(5)}, 2, str(out), seed=0)
    planner.close()
    with open(out) as f:
        got = list(csv.DictReader(f))
    assert len(got) == 2 and list(got[0].keys()) == EXPECTED
    assert all(1 <= int(r["len"]) <= 50 and float(r["num_sims"]) == 64.0 for r in got)
    assert rows[0]["len"] == int(got[0]["len"])
