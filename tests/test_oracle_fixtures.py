"""The CPU oracle against every fixture the reference tree holds for the path
(SURVEY.md section 8(c)), distilled into tests/golden/reference_fixtures.json, and
against published known-answer vectors.  CPU only."""
import random
from collections import Counter, deque

import numpy as np
import pytest

from tests.conftest import golden


def test_philox_known_answers():
    # Random123 kat_vectors: philox4x32 10 rounds
    from oracle.philox import philox4x32_10

    def run(c, k):
        return [int(x[0]) for x in philox4x32_10(*[np.array([v]) for v in c], *k)]

    assert run((0, 0, 0, 0), (0, 0)) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert run((0xFFFFFFFF,) * 4, (0xFFFFFFFF, 0xFFFFFFFF)) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert run((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0)) == [
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


@pytest.fixture(scope="module")
def driving(oracle_env):
    from oracle.driving import DrivingModel

    return DrivingModel()


def test_driving_obs_width_and_actions(driving, reference_fixtures):
    from gymnasium import spaces

    fx = reference_fixtures["Driving-v1"]
    assert spaces.flatdim(driving.observation_spaces["0"]) == fx["obs_dim"] == 124
    assert driving.action_spaces["0"].n == fx["num_actions"] == 5


def test_driving_active_inputs_match_checkpoints(driving, reference_fixtures):
    """One-hot inputs the oracle can ever produce under play == inputs that ever
    received gradient while the shipped BR policies trained (Adam exp_avg_sq != 0),
    except `crashed=1` observations, which training never sees (auto-reset)."""
    from gymnasium import spaces

    fx = reference_fixtures["Driving-v1"]["active_inputs"]
    sp = driving.observation_spaces["0"]
    seen = np.zeros(124, bool)
    rng = random.Random(0)
    driving.seed(0)
    for _ in range(600):
        s = driving.sample_initial_state()
        o = driving.sample_initial_obs(s)
        for t in range(50):
            for i in driving.possible_agents:
                if not s[int(i)].crashed:
                    seen |= spaces.flatten(sp, o[i]).astype(bool)
            # half random, half "drive forward" so the vehicles actually get around
            a = {i: (rng.randrange(5) if rng.random() < 0.6 else 1) for i in driving.possible_agents}
            ts = driving.step(s, a)
            s, o = ts.state, ts.observations
            if ts.all_done:
                break
    expected = np.array([c == "1" for c in fx])
    # the oracle never produces an input the reference never produced
    assert not (seen & ~expected).any(), np.nonzero(seen & ~expected)
    # structural columns: ego cell, speed, coords, dest coords, flags
    assert list(np.nonzero(expected[40:44])[0]) == [0]          # ego cell (index 10) is always VEHICLE
    assert seen[40] and not seen[41:44].any()
    assert (seen[92:120] == expected[92:120]).all()             # destination one-hots: x,y in {0,1,12}
    assert expected[64:92].all()                                # current coord: every x and y occurs
    assert (seen[:40] | ~expected[:40]).sum() >= 38             # local cells (rare ones may not recur)


def test_driving_layout_distance_histogram(driving, reference_fixtures):
    """16 (start, dest) shortest-path distances are {22:2, 23:4, 24:2, 25:4, 26:4}: the
    D0 classes recovered from the float32 residuals of the successful returns appear in
    ratio ~ 1:2:1:2:2."""
    g = driving.grid
    hist = Counter(g.get_shortest_path_distance(s, d) for s in g.start_coords[0] for d in g.dest_coords[0])
    assert dict(hist) == {22: 2, 23: 4, 24: 2, 25: 4, 26: 4}
    fx = reference_fixtures["Driving-v1"]["rewards"]["success_residual_hist_2p32"]
    f32 = lambda x: float(np.float32(x))  # noqa: E731
    by_d0 = Counter()
    for d0 in range(22, 27):
        for kf in (1, 2):
            # whole-episode float32-rounded reward sum - 1, in units of 2^-32, depends only on (D0, kf)
            n1 = d0 - 2 * 0 - (kf == 1) - 2 * (kf == 2)
            tot = n1 * f32(1 / (2 * d0)) + f32(0.5 + kf / (2 * d0))
            key = str(int(round((tot - 1) * 2**32)))
            assert key in fx, (d0, kf, key)
            by_d0[d0] += fx[key]
    assert len(fx) == 10
    total = sum(by_d0.values())
    for d0, w in hist.items():
        assert abs(by_d0[d0] / total - w / 16) < 0.015


def test_driving_min_time_to_destination(driving, reference_fixtures):
    """Single-vehicle dynamics: fastest possible arrival per D0 class equals the shortest
    successful episode of the reference's BR policies in that class."""
    from oracle.driving import ACTIONS, STOPPED

    m, g = driving, driving.grid

    def mintime(start, dest):
        s0 = (start, 0, STOPPED)
        seen = {s0: 0}
        q = deque([s0])
        while q:
            cur = q.popleft()
            c, d, sp = cur
            if c == dest:
                return seen[cur]
            for a in ACTIONS:
                nsp = m._get_next_speed(a, sp)
                md = m._get_move_direction(a, nsp, d)
                nd = d if nsp == 0 else m._get_next_direction(a, d)
                nc, _, _ = m._get_next_coord(c, nsp, md, set())
                ns = (nc, nd, nsp)
                if ns not in seen:
                    seen[ns] = seen[cur] + 1
                    q.append(ns)

    best = {}
    for s in g.start_coords[0]:
        for d in g.dest_coords[0]:
            d0 = g.get_shortest_path_distance(s, d)
            best[d0] = min(best.get(d0, 99), mintime(s, d))
    fx = reference_fixtures["Driving-v1"]["rewards"]
    f32 = lambda x: float(np.float32(x))  # noqa: E731
    ref_min = {}
    for d0 in range(22, 27):
        for kf in (1, 2):
            tot = (d0 - kf) * f32(1 / (2 * d0)) + f32(0.5 + kf / (2 * d0))
            key = str(int(round((tot - 1) * 2**32)))
            ref_min[d0] = min(ref_min.get(d0, 99), fx["success_min_len_by_residual"][key])
    assert best == ref_min == {22: 15, 23: 15, 24: 16, 25: 16, 26: 17}


def test_driving_random_play_statistics(driving, reference_fixtures):
    """Uniform-random joint play vs div_results.csv (100 episodes there; loose bounds)."""
    fx = reference_fixtures["Driving-v1"]
    rng = random.Random(1)
    driving.seed(1)
    n, crashes, wins, rets = 1500, 0, 0, []
    for _ in range(n):
        s = driving.sample_initial_state()
        r = 0.0
        for t in range(50):
            ts = driving.step(s, {i: rng.randrange(5) for i in driving.possible_agents})
            s = ts.state
            r += ts.rewards["0"]
            if ts.all_done:
                break
        crashes += s[0].crashed
        wins += s[0].dest_reached
        rets.append(r)
    ref = fx["random_vs_random"][0]
    se = ref["episode_reward_std"] / 10
    assert wins / n < 0.01 and ref["num_WIN"] == 0
    assert abs(np.mean(rets) - ref["episode_reward_mean"]) < 3 * se
    assert abs(crashes / n - ref["num_LOSS"] / 100) < 3 * 0.043
    # a lone random vehicle: progress reward only
    from oracle.driving import DrivingModel

    solo = DrivingModel(num_agents=1)
    solo.seed(2)
    prog = []
    for _ in range(3000):
        s = solo.sample_initial_state()
        r = 0.0
        for t in range(50):
            ts = solo.step(s, {"0": rng.randrange(5)})
            s = ts.state
            r += ts.rewards["0"]
        prog.append(r)
    assert abs(np.mean(prog) - fx["random_uncrashed_reward_mean"][0]) < 0.012


def test_driving_reward_lattice(driving, reference_fixtures):
    """Failure returns are -1 + k/(2 D0) or k/(2 D0): denominators divide 2*D0, D0 in 22..26."""
    fx = reference_fixtures["Driving-v1"]["rewards"]["failure_return_denominators"]
    for den in fx:
        assert any((2 * d0) % int(den) == 0 for d0 in range(22, 27)), den
    # and the oracle's rewards live on the same lattice
    rng = random.Random(3)
    driving.seed(3)
    for _ in range(200):
        s = driving.sample_initial_state()
        d0 = s[0].init_dest_dist
        tot = 0.0
        for t in range(50):
            ts = driving.step(s, {i: rng.randrange(5) for i in driving.possible_agents})
            s = ts.state
            tot += ts.rewards["0"]
            if ts.all_done:
                break
        base = tot + 1 if s[0].crashed else tot
        assert abs(base * 2 * d0 - round(base * 2 * d0)) < 1e-9


def test_oracle_reproduces_golden(driving):
    """Regression pin: the committed golden vectors are what the oracle computes."""
    from tests.golden.make_golden import obs_fields, vs_fields
    from oracle.driving import VehicleState

    g = golden("driving_step.npz")
    idx = np.random.default_rng(0).choice(len(g["state"]), 600, replace=False)
    for i in idx:
        st = tuple(VehicleState((int(f[0]), int(f[1])), int(f[2]), int(f[3]), (int(f[4]), int(f[5])),
                                int(f[6]), int(f[7]), int(f[8]), int(f[9])) for f in g["state"][i])
        ts = driving.step_with_noise(st, {"0": int(g["action"][i, 0]), "1": int(g["action"][i, 1])}, int(g["order"][i]))
        assert [vs_fields(v) for v in ts.state] == g["next_state"][i].tolist()
        assert [obs_fields(ts.observations[a]) for a in ("0", "1")] == g["obs"][i].tolist()
        assert [ts.rewards[a] for a in ("0", "1")] == g["reward"][i].tolist()
        assert ts.all_done == bool(g["all_done"][i])
    # the golden set exercises the interesting branches
    assert (g["reward"] == -1.0).any() and (g["reward"] > 0.5).any() and g["all_done"].any()
    assert ((g["reward"] < 0) & (g["reward"] > -1)).any()   # crash with progress in the same step


# ---------------------------------------------------------------- LevelBasedForaging-v3
@pytest.fixture(scope="module")
def lbf(oracle_env):
    from oracle.lbf import LevelBasedForagingModel

    return LevelBasedForagingModel()


def test_lbf_obs_width_and_actions(lbf, reference_fixtures):
    from gymnasium import spaces

    fx = reference_fixtures["LevelBasedForaging-v3"]
    assert spaces.flatdim(lbf.observation_spaces["0"]) == fx["obs_dim"] == 300
    assert lbf.action_spaces["0"].n == fx["num_actions"] == 6
    assert fx["env_kwargs"]["static_layout"] is True and fx["env_kwargs"]["observation_mode"] == "tuple"


def test_lbf_active_inputs_and_random_play(lbf, reference_fixtures):
    """(1) the oracle never fires a one-hot input that never received gradient while the
    shipped BR policies trained (Adam exp_avg_sq mask); (2) uniform-random joint play matches
    div_results.csv (100 episodes there: 0.0597 +- 0.088 / 0.0158 +- 0.041)."""
    from gymnasium import spaces

    fx = reference_fixtures["LevelBasedForaging-v3"]
    expected = np.array([c == "1" for c in fx["active_inputs"]])
    sp = lbf.observation_spaces["0"]
    seen = np.zeros(300, bool)
    rng = random.Random(0)
    lbf.seed(0)
    rets = [[], []]
    for _ in range(1500):
        s = lbf.sample_initial_state()
        o = lbf.sample_initial_obs(s)
        r = [0.0, 0.0]
        for t in range(50):
            for i in lbf.possible_agents:
                seen |= spaces.flatten(sp, o[i]).astype(bool)
            ts = lbf.step(s, {i: rng.randrange(6) for i in lbf.possible_agents})
            s, o = ts.state, ts.observations
            r[0] += ts.rewards["0"]; r[1] += ts.rewards["1"]
            if ts.all_done:
                break
        rets[0].append(r[0]); rets[1].append(r[1])
    assert not (seen & ~expected).any(), np.nonzero(seen & ~expected)
    # structure of the mask: self x / y one-hots fire only at value `sight` (index sight + 1)
    assert list(np.nonzero(expected[0:11])[0]) == [3] and list(np.nonzero(expected[11:22])[0]) == [3]
    assert (seen & expected).sum() >= 0.9 * expected.sum()
    for k, ref in enumerate(fx["random_vs_random"]):
        se = ref["episode_reward_std"] / 10
        assert abs(np.mean(rets[k]) - ref["episode_reward_mean"]) < 3 * se, (k, np.mean(rets[k]))


def test_lbf_reward_lattice(lbf, reference_fixtures):
    """Episode returns are sums of level*food_level/(loader level sum * spawned): rationals in
    [0, 1]; both agents together never collect more than 1 per episode, exactly 1 when every
    food is loaded."""
    from fractions import Fraction

    from tests.golden.make_golden import lbf_greedy_action

    fx = reference_fixtures["LevelBasedForaging-v3"]["lattice"]
    assert fx["max_len"] == 50
    assert all(0 < Fraction(r) <= 1 for r, _ in fx["top_returns"])
    rng = random.Random(4)
    lbf.seed(4)
    finished = 0
    for _ in range(150):
        s = lbf.sample_initial_state()
        tot = Fraction(0)
        for t in range(200):
            a = {i: lbf_greedy_action(lbf, s, int(i), rng) for i in lbf.possible_agents}
            before = s
            ts = lbf.step(s, a)
            s = ts.state
            for i in lbf.possible_agents:
                if ts.rewards[i]:
                    tot += Fraction(ts.rewards[i]).limit_denominator(10000)
            assert s.food_spawned == before.food_spawned
            if ts.all_done:
                finished += 1
                assert tot == 1
                break
        assert tot <= 1
    assert finished > 20


def test_lbf_oracle_reproduces_golden(lbf):
    from tests.golden.make_golden import lbf_fields, lbf_state

    g = golden("lbf_step.npz")
    idx = np.random.default_rng(0).choice(len(g["state"]), 800, replace=False)
    for i in idx:
        ts = lbf.step(lbf_state(lbf, g["state"][i]), {"0": int(g["action"][i, 0]), "1": int(g["action"][i, 1])})
        assert lbf_fields(ts.state) == g["next_state"][i].tolist()
        assert [list(ts.observations[a]) for a in ("0", "1")] == g["obs"][i].tolist()
        assert [ts.rewards[a] for a in ("0", "1")] == g["reward"][i].tolist()
        assert ts.all_done == bool(g["all_done"][i])
