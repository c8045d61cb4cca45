"""How the oracles' last open details were found: the reference ships best-response policies trained on the REAL
environments (baseline_exps/env_data/<env>/br_models); a candidate restatement is scored by the return two of them
collect in self-play on it.  CPU only, needs /root/reference (build container):

    python tools/policy_probe.py pp          # PredatorPrey: 8 window symmetries x 6 cell-code permutations
    python tools/policy_probe.py lbf         # LevelBasedForaging: food placements, player corners, (x, y) order
    python tools/policy_probe.py lbf-random  # uniform-random returns per placement against div_results.csv
    python tools/policy_probe.py lbf-agent0  # the trained agent 0 alone: return vs a random partner, critics' V(o0)

Results of the committed oracles are asserted in tests/test_oracle_policies.py; the numbers are in DESIGN.md section 3.
"""
from __future__ import annotations

import itertools
import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import shims  # noqa: E402

shims.install()
from tests.test_oracle_policies import _policies, _self_play  # noqa: E402

torch.set_num_threads(1)
LATTICE = [1, 3, 5, 7]


def probe_pp(episodes=12):
    from oracle.predator_prey import PredatorPreyModel

    nets = _policies("PredatorPrey-v0", 5)
    idx = np.arange(25).reshape(5, 5)
    rows = []
    for transpose, flip_a, flip_b in itertools.product((0, 1), repeat=3):
        a = idx.T if transpose else idx
        a = a[::-1, :] if flip_a else a
        a = a[:, ::-1] if flip_b else a
        cells = a.reshape(-1)
        for codes in itertools.permutations([0, 1, 3]):          # EMPTY, WALL, PREY relabelled; PREDATOR is pinned
            remap = {0: codes[0], 1: codes[1], 2: 2, 3: codes[2]}
            ret = _self_play(PredatorPreyModel(), nets, episodes, seed=0,
                             obs_map=lambda o: tuple(remap[o[k]] for k in cells))
            rows.append((round(float(ret[:, 0].mean()), 3), ("transposed" if transpose else "as is", flip_a, flip_b), codes))
            print(rows[-1], flush=True)
    rows.sort(reverse=True)
    print("best:", rows[:5])


def _lbf(food, players):
    from oracle.lbf import LevelBasedForagingModel

    m = LevelBasedForagingModel()
    m.food_coords, m.player_coords = list(food), list(players)
    return m


def probe_lbf(episodes=16):
    nets = _policies("LevelBasedForaging-v3", 6)
    foods = {}
    for a in range(0, 8):
        for r0 in range(0, 4):
            rows = [r0 + 2 * k for k in range(4)]
            if rows[-1] <= 9:
                foods[f"columns {a},{a + 2} rows {rows}"] = [(x, y) for x in (a, a + 2) for y in rows]
                foods[f"rows {a},{a + 2} columns {rows}"] = [(x, y) for y in (a, a + 2) for x in rows]
    for ya, yb in ((1, 5), (3, 7), (1, 3), (5, 7), (3, 5), (1, 7)):
        foods[f"four columns, rows {ya},{yb}"] = [(x, y) for x in LATTICE for y in (ya, yb)]
    foods["checkerboard A"] = [(x, y) for x in LATTICE for y in LATTICE if ((x + y) // 2) % 2 == 0]
    foods["checkerboard B"] = [(x, y) for x in LATTICE for y in LATTICE if ((x + y) // 2) % 2 == 1]
    corners = {"(0,9) (9,0)": [(0, 9), (9, 0)], "(0,0) (9,9)": [(0, 0), (9, 9)], "(9,0) (0,9)": [(9, 0), (0, 9)],
               "(9,9) (0,0)": [(9, 9), (0, 0)]}
    swap = lambda o: tuple(o[k + (1 if k % 3 == 0 else -1 if k % 3 == 1 else 0)] for k in range(len(o)))  # noqa: E731
    out = []
    for name, cells in foods.items():
        for cname, players in (corners.items() if name.startswith("columns 5,7 rows [1") or name.startswith("columns 1,3 rows [1") else
                               [("(0,9) (9,0)", corners["(0,9) (9,0)"])]):
            ret = _self_play(_lbf(cells, players), nets, episodes, seed=0).mean(0)
            out.append((round(float(ret.sum()), 3), name, cname, ret.round(3).tolist()))
            print(out[-1], flush=True)
    out.sort(reverse=True)
    print("best:", out[:6])
    best = [(x, y) for x in (5, 7) for y in LATTICE]
    print("best placement with (x, y) swapped in the observation:",
          _self_play(_lbf(best, [(9, 0), (0, 9)]), nets, episodes, seed=0, obs_map=swap).mean(0).round(3))


def probe_lbf_random(episodes=1500):
    print("div_results.csv, Random-v0 vs Random-v0: 0.0597 +- 0.0088 (agent 0), 0.0158 +- 0.0041 (agent 1)")
    for cols in ((1, 3), (5, 7)):
        for players in ([(0, 9), (9, 0)], [(0, 0), (9, 9)], [(9, 0), (0, 9)], [(9, 9), (0, 0)]):
            m = _lbf([(x, y) for x in cols for y in LATTICE], players)
            m.seed(0)
            rng = random.Random(0)
            tot = np.zeros((episodes, 2))
            for e in range(episodes):
                s = m.sample_initial_state()
                for _ in range(50):
                    ts = m.step(s, {"0": rng.randrange(6), "1": rng.randrange(6)})
                    s = ts.state
                    tot[e] += [ts.rewards["0"], ts.rewards["1"]]
                    if ts.all_done:
                        break
            print("food columns", cols, "players", players, tot.mean(0).round(4))


def probe_lbf_agent0(episodes=80):
    """The two placement families the fixtures leave open (oracle/lbf.py header), by what the trained AGENT 0 does on
    them: its return against a uniform-random partner (the role and start it was trained in; no second copy playing a
    role it never saw) and the critics' value of the first observation, over all ten checkpoints."""
    from gymnasium import spaces

    names = [f"P{p}_seed{k}.pt" for p in (0, 1) for k in range(5)]
    nets = [_policies("LevelBasedForaging-v3", 6, (n,))[0] for n in names]
    even = [2, 4, 6, 8]
    cands = {"adopted: foods x in {5,7}, agents (9,0) / (0,9)": ([(x, y) for x in (5, 7) for y in LATTICE], [(9, 0), (0, 9)]),
             "round 1: foods x in {1,3}, agents (0,9) / (9,0)": ([(x, y) for x in (1, 3) for y in LATTICE], [(0, 9), (9, 0)]),
             "second family: foods x in {1,3}, y even, agents (0,0) / (9,9)": ([(x, y) for x in (1, 3) for y in even], [(0, 0), (9, 9)]),
             "second family: foods x in {2,4}, agents (0,0) / (9,9)": ([(x, y) for x in (2, 4) for y in LATTICE], [(0, 0), (9, 9)]),
             "foods out of sight: x in {5,7}, agents (0,9) / (9,0)": ([(x, y) for x in (5, 7) for y in LATTICE], [(0, 9), (9, 0)])}
    for name, (food, players) in cands.items():
        m = _lbf(food, players)
        space = m.observation_spaces["0"]
        m.seed(0)
        torch.manual_seed(0)
        rng = random.Random(0)
        total, values = 0.0, []
        with torch.no_grad():
            for ep in range(episodes):
                net = nets[ep % 2]                       # the two checkpoints of the tests
                s = m.sample_initial_state()
                obs = m.sample_initial_obs(s)
                x0 = torch.tensor(spaces.flatten(space, obs["0"]), dtype=torch.float32)
                for critic_net in nets:
                    out, _ = critic_net.lstm(critic_net.trunk(x0).view(1, 1, -1), (torch.zeros(1, 1, 64), torch.zeros(1, 1, 64)))
                    values.append(float(critic_net.critic(out.view(1, -1))))
                hidden = (torch.zeros(1, 1, 64), torch.zeros(1, 1, 64))
                for _ in range(50):
                    x = torch.tensor(spaces.flatten(space, obs["0"]), dtype=torch.float32)
                    a, hidden = net.act(x, hidden)
                    ts = m.step(s, {"0": a, "1": rng.randrange(6)})
                    s, obs = ts.state, ts.observations
                    total += ts.rewards["0"]
                    if ts.all_done:
                        break
        print(f"{name}: agent 0 vs random partner {total / episodes:.3f}, critics' V(first observation) {np.mean(values):.3f}", flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "pp"
    {"pp": probe_pp, "lbf": probe_lbf, "lbf-random": probe_lbf_random, "lbf-agent0": probe_lbf_agent0}[what]()
