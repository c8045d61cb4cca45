"""Generate the committed golden vectors from the CPU oracle (run in the build
container: `python tests/golden/make_golden.py`).  Representation-independent: raw
field arrays, not the kernels' packed words."""
from __future__ import annotations

import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import shims  # noqa: E402

shims.install()
from oracle import rollout as R  # noqa: E402
from oracle.driving import DrivingModel, VehicleState  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def vs_fields(vs):
    return [vs.coord[0], vs.coord[1], vs.facing_dir, vs.speed, vs.dest_coord[0], vs.dest_coord[1],
            vs.dest_reached, vs.crashed, vs.min_dest_dist, vs.init_dest_dist]


def obs_fields(o):
    cells, speed, coord, dest, reached, crashed = o
    return [*cells, speed, coord[0], coord[1], dest[0], dest[1], reached, crashed]


def synthetic_state(m, rng):
    """Two vehicles on distinct road cells with random heading / speed (dense in
    near-collision configurations, which random play reaches rarely)."""
    g = m.grid
    road = g.unblocked_coords
    dests = g.dest_coords[0]
    c0 = rng.choice(road)
    near = [c for c in road if c != c0 and g.manhattan_dist(c, c0) <= 3]
    c1 = rng.choice(near) if near and rng.random() < 0.7 else rng.choice([c for c in road if c != c0])
    d = rng.sample(dests, 2)
    out = []
    for c, dest in zip((c0, c1), d):
        dist = g.get_shortest_path_distance(c, dest)
        mind = max(0, dist - rng.randrange(3))
        d0 = min(63, max(1, mind + rng.randrange(8)))
        crashed = int(rng.random() < 0.05)
        out.append(VehicleState(c, rng.randrange(4), rng.randrange(4), dest, int(c == dest), crashed, mind, d0))
    return tuple(out)


def driving_step_cases(n_play=2500, n_syn=2500, seed=20261018):
    m = DrivingModel()
    m.seed(seed)
    rng = random.Random(seed)
    states, actions, orders = [], [], []
    while len(states) < n_play:
        s = m.sample_initial_state()
        for _ in range(51):  # one step past the horizon: stepping a terminal state is legal
            a = [rng.randrange(5), rng.randrange(5)]
            o = rng.randrange(2)
            states.append(s); actions.append(a); orders.append(o)
            ts = m.step_with_noise(s, {"0": a[0], "1": a[1]}, o)
            s = ts.state
            if ts.all_done and rng.random() < 0.7:
                break
    for _ in range(n_syn):
        states.append(synthetic_state(m, rng))
        actions.append([rng.randrange(5), rng.randrange(5)])
        orders.append(rng.randrange(2))
    S, S2, O, RW, T, AD = [], [], [], [], [], []
    for s, a, o in zip(states, actions, orders):
        ts = m.step_with_noise(s, {"0": a[0], "1": a[1]}, o)
        S.append([vs_fields(v) for v in s])
        S2.append([vs_fields(v) for v in ts.state])
        O.append([obs_fields(ts.observations[i]) for i in m.possible_agents])
        RW.append([ts.rewards[i] for i in m.possible_agents])
        T.append([ts.terminations[i] for i in m.possible_agents])
        AD.append(ts.all_done)
    return dict(
        state=np.array(S, np.int16), action=np.array(actions, np.uint8), order=np.array(orders, np.uint8),
        next_state=np.array(S2, np.int16), obs=np.array(O, np.int8), reward=np.array(RW, np.float64),
        terminated=np.array(T, np.bool_), all_done=np.array(AD, np.bool_),
    )


def driving_init_cases():
    """Every legal pick combination of sample_initial_state + its initial obs."""
    m = DrivingModel()
    picks, S, O = [], [], []
    for a in range(4):
        for b in range(4):
            for c in range(3):
                for d in range(3):
                    p = [a, b, c, d]
                    s = m.initial_state_from_noise(p)
                    o = m.sample_initial_obs(s)
                    picks.append(p); S.append([vs_fields(v) for v in s])
                    O.append([obs_fields(o[i]) for i in m.possible_agents])
    return dict(picks=np.array(picks, np.uint8), state=np.array(S, np.int16), obs=np.array(O, np.int8))


def driving_rollout_cases(n=512, seed=7, discount=0.95, max_steps=50):
    """Philox-driven rollouts: initial state from (seed, particle), random joint actions."""
    m = DrivingModel()
    particles = np.arange(n, dtype=np.uint64) + np.uint64(1000)
    picks = R.driving_initial_picks(m, seed, particles)
    rets, steps, finals, inits = [], [], [], []
    for p, pk in zip(particles, picks):
        s0 = m.initial_state_from_noise(pk)
        ret, k, s = R.rollout(m, s0, "0", max_steps, discount, seed, int(p))
        rets.append(ret); steps.append(k)
        inits.append([vs_fields(v) for v in s0]); finals.append([vs_fields(v) for v in s])
    return dict(seed=np.array(seed), particle=particles, discount=np.array(discount), max_steps=np.array(max_steps),
                init_state=np.array(inits, np.int16), ret=np.array(rets, np.float64),
                steps=np.array(steps, np.uint8), final_state=np.array(finals, np.int16))


# ---------------------------------------------------------------- PredatorPrey-v0
def pp_fields(s):
    return [*[v for c in s.predator_coords for v in c], *[v for c in s.prey_coords for v in c], *s.prey_caught]


def pp_step_cases(n_play=2500, n_syn=2500, seed=20261019):
    from oracle.predator_prey import PPState, PredatorPreyModel

    m = PredatorPreyModel()
    m.seed(seed)
    rng = random.Random(seed)
    states = []
    while len(states) < n_play:
        s = m.sample_initial_state()
        for _ in range(51):
            states.append(s)
            # chase the prey a bit so that captures and crowded cells occur
            acts = {i: rng.randrange(5) for i in m.possible_agents}
            ts = m.step(s, acts)
            s = ts.state
            if ts.all_done and rng.random() < 0.7:
                break
    cells = [(x, y) for x in range(10) for y in range(10)]
    for _ in range(n_syn):
        c0 = rng.choice(cells)
        near = [c for c in cells if c != c0 and abs(c[0] - c0[0]) + abs(c[1] - c0[1]) <= 3]
        picks = [c0] + rng.sample(near, 4) if rng.random() < 0.8 else rng.sample(cells, 5)
        rng.shuffle(picks)
        caught = tuple(int(rng.random() < 0.2) for _ in range(3))
        states.append(PPState(tuple(picks[:2]), tuple(picks[2:]), caught))
    S, A, W, S2, O, RW, AD = [], [], [], [], [], [], []
    for s in states:
        a = [rng.randrange(5), rng.randrange(5)]
        w = [rng.getrandbits(32) for _ in range(3)]
        ts = m.step_with_noise(s, {"0": a[0], "1": a[1]}, w)
        S.append(pp_fields(s)); A.append(a); W.append(w); S2.append(pp_fields(ts.state))
        O.append([list(ts.observations[i]) for i in m.possible_agents])
        RW.append([ts.rewards[i] for i in m.possible_agents]); AD.append(ts.all_done)
    return dict(state=np.array(S, np.int8), action=np.array(A, np.uint8), noise=np.array(W, np.uint32),
                next_state=np.array(S2, np.int8), obs=np.array(O, np.int8), reward=np.array(RW, np.float64),
                all_done=np.array(AD, np.bool_))


def pp_init_cases(n=400, seed=5):
    from oracle.predator_prey import PredatorPreyModel

    m = PredatorPreyModel()
    rng = random.Random(seed)
    picks, S, O = [], [], []
    for _ in range(n):
        p = [rng.randrange(b) for b in m.initial_noise_bounds()]
        s = m.initial_state_from_noise(p)
        o = m.sample_initial_obs(s)
        picks.append(p); S.append(pp_fields(s)); O.append([list(o[i]) for i in m.possible_agents])
    return dict(picks=np.array(picks, np.uint8), state=np.array(S, np.int8), obs=np.array(O, np.int8))


def pp_rollout_cases(n=384, seed=9, discount=0.95, max_steps=50):
    from oracle.predator_prey import PredatorPreyModel

    m = PredatorPreyModel()
    particles = np.arange(n, dtype=np.uint64) + np.uint64(777)
    picks = R.pp_initial_picks(m, seed, particles)
    rets, steps, finals, inits = [], [], [], []
    for p, pk in zip(particles, picks):
        s0 = m.initial_state_from_noise([int(v) for v in pk])
        ret, k, s = R.pp_rollout(m, s0, "1", max_steps, discount, seed, int(p))
        rets.append(ret); steps.append(k); inits.append(pp_fields(s0)); finals.append(pp_fields(s))
    return dict(seed=np.array(seed), particle=particles, discount=np.array(discount), max_steps=np.array(max_steps),
                ego=np.array(1), init_state=np.array(inits, np.int8), ret=np.array(rets, np.float64),
                steps=np.array(steps, np.uint8), final_state=np.array(finals, np.int8))

# ---------------------------------------------------------------- LevelBasedForaging-v3
def lbf_fields(s):
    """[p0 x, y, level, p1 x, y, level, 8 food levels, food_spawned] (food cells are static)."""
    return [*[v for p in s.players for v in (p.coord[0], p.coord[1], p.level)], *[f.level for f in s.food],
            s.food_spawned]


def lbf_state(m, f):
    from oracle.lbf import Food, LBFState, Player

    f = [int(v) for v in f]
    players = (Player(0, (f[0], f[1]), f[2]), Player(1, (f[3], f[4]), f[5]))
    food = tuple(Food(m.food_coords[k], f[6 + k]) for k in range(8))
    return LBFState(players, food, f[14])


def lbf_greedy_action(m, s, idx, rng):
    """Walk towards the nearest remaining food and LOAD next to it (random play almost never
    loads; this reaches captures, joint loads, failed loads and the terminal state)."""
    from oracle import lbf as L

    p = s.players[idx]
    foods = [f for f in s.food if f.level > 0]
    if not foods or rng.random() < 0.25:
        return rng.randrange(6)
    f = min(foods, key=lambda f: abs(f.coord[0] - p.coord[0]) + abs(f.coord[1] - p.coord[1]))
    dx, dy = f.coord[0] - p.coord[0], f.coord[1] - p.coord[1]
    if abs(dx) + abs(dy) == 1:
        return L.LOAD
    if abs(dx) >= abs(dy) and dx != 0:
        return L.EAST if dx > 0 else L.WEST
    return L.SOUTH if dy > 0 else L.NORTH


def lbf_step_cases(n_play=3000, n_syn=3000, seed=20261020):
    from oracle.lbf import Food, LBFState, LevelBasedForagingModel, Player

    m = LevelBasedForagingModel()
    m.seed(seed)
    rng = random.Random(seed)
    states, actions = [], []
    while len(states) < n_play:
        s = m.sample_initial_state()
        greedy = rng.random() < 0.7
        for _ in range(51):
            a = [lbf_greedy_action(m, s, i, rng) if greedy else rng.randrange(6) for i in range(2)]
            states.append(s); actions.append(a)
            ts = m.step(s, {"0": a[0], "1": a[1]})
            s = ts.state
            if ts.all_done and rng.random() < 0.7:
                break
    cells = [(x, y) for x in range(m.size) for y in range(m.size)]
    free = [c for c in cells if c not in m.food_coords]
    for _ in range(n_syn):
        # players next to food / next to each other, random levels, some food already collected
        fc = rng.choice(m.food_coords)
        near = [c for c in free if abs(c[0] - fc[0]) + abs(c[1] - fc[1]) <= 2]
        c0 = rng.choice(near) if rng.random() < 0.8 else rng.choice(free)
        near1 = [c for c in free if c != c0 and abs(c[0] - c0[0]) + abs(c[1] - c0[1]) <= 2]
        c1 = rng.choice(near1) if near1 and rng.random() < 0.8 else rng.choice([c for c in free if c != c0])
        lv = [rng.randrange(1, 4), rng.randrange(1, 4)]
        food = [Food(c, rng.randrange(1, lv[0] + lv[1] + 1)) for c in m.food_coords]
        spawned = sum(f.level for f in food)
        left = rng.choice([1, 1, 2, 8, 8, 8])
        keep = set(rng.sample(range(8), left)) | {m.food_coords.index(fc)}
        food = tuple(f if k in keep else Food(f.coord, 0) for k, f in enumerate(food))
        states.append(LBFState((Player(0, c0, lv[0]), Player(1, c1, lv[1])), food, spawned))
        actions.append([rng.choice([5, 5, rng.randrange(6)]), rng.choice([5, 5, rng.randrange(6)])])
    S, S2, O, RW, AD = [], [], [], [], []
    for s, a in zip(states, actions):
        ts = m.step(s, {"0": a[0], "1": a[1]})
        S.append(lbf_fields(s)); S2.append(lbf_fields(ts.state))
        O.append([list(ts.observations[i]) for i in m.possible_agents])
        RW.append([ts.rewards[i] for i in m.possible_agents]); AD.append(ts.all_done)
    return dict(state=np.array(S, np.int8), action=np.array(actions, np.uint8), next_state=np.array(S2, np.int8),
                obs=np.array(O, np.int8), reward=np.array(RW, np.float64), all_done=np.array(AD, np.bool_))


def lbf_init_cases(n=600, seed=6):
    from oracle.lbf import LevelBasedForagingModel

    m = LevelBasedForagingModel()
    rng = random.Random(seed)
    picks, S, O = [], [], []
    for _ in range(n):
        p = []
        for _ in range(10):
            p.append(rng.randrange(m.initial_noise_bounds(p)))
        s = m.initial_state_from_noise(p)
        o = m.sample_initial_obs(s)
        picks.append(p); S.append(lbf_fields(s)); O.append([list(o[i]) for i in m.possible_agents])
    return dict(picks=np.array(picks, np.uint8), state=np.array(S, np.int8), obs=np.array(O, np.int8))


def lbf_rollout_cases(n=384, seed=11, discount=0.95, max_steps=50):
    from oracle.lbf import LevelBasedForagingModel

    m = LevelBasedForagingModel()
    particles = np.arange(n, dtype=np.uint64) + np.uint64(4242)
    picks = R.lbf_initial_picks(m, seed, particles)
    rets, steps, finals, inits = [], [], [], []
    for p, pk in zip(particles, picks):
        s0 = m.initial_state_from_noise([int(v) for v in pk])
        ret, k, s = R.lbf_rollout(m, s0, "0", max_steps, discount, seed, int(p))
        rets.append(ret); steps.append(k); inits.append(lbf_fields(s0)); finals.append(lbf_fields(s))
    return dict(seed=np.array(seed), particle=particles, discount=np.array(discount), max_steps=np.array(max_steps),
                ego=np.array(0), init_state=np.array(inits, np.int8), ret=np.array(rets, np.float64),
                steps=np.array(steps, np.uint8), final_state=np.array(finals, np.int8))


if __name__ == "__main__":
    only = sys.argv[1:]
    if not only or "lbf" in only:
        np.savez_compressed(os.path.join(HERE, "lbf_step.npz"), **lbf_step_cases())
        np.savez_compressed(os.path.join(HERE, "lbf_init.npz"), **lbf_init_cases())
        np.savez_compressed(os.path.join(HERE, "lbf_rollout.npz"), **lbf_rollout_cases())
    if not only or "pp" in only:
        np.savez_compressed(os.path.join(HERE, "pp_step.npz"), **pp_step_cases())
        np.savez_compressed(os.path.join(HERE, "pp_init.npz"), **pp_init_cases())
        np.savez_compressed(os.path.join(HERE, "pp_rollout.npz"), **pp_rollout_cases())
    if only:
        sys.exit(0)
    np.savez_compressed(os.path.join(HERE, "driving_step.npz"), **driving_step_cases())
    np.savez_compressed(os.path.join(HERE, "driving_init.npz"), **driving_init_cases())
    np.savez_compressed(os.path.join(HERE, "driving_rollout.npz"), **driving_rollout_cases())
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
