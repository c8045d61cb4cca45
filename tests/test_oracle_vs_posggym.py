"""Cross-check of the CPU oracle against the REAL upstream posggym models -- runs only where
``import posggym`` resolves to the real package (it is absent from the build container and from the GPU
box, where this module is skipped: oracle parity against upstream stays "unpinned" until it has run
somewhere).  SURVEY.md section 7 "Hard parts": run it automatically whenever posggym is importable.

Alignment without sharing an RNG stream: the real model draws its step noise from ``model.rng``; the
oracle takes the noise as an argument.  Every real step is compared with the oracle's outcomes over ALL
values of that noise (Driving: the 2 processing orders; PredatorPrey: every per-prey choice index;
LevelBasedForaging: none) through the model's public outputs -- observations, rewards,
terminations, ``all_done`` -- plus, where the observations hide part of the state (prey out of sight), the
converted next state; the oracle continues from the matching outcome.  No candidate matching = the oracle
disagrees with upstream.
"""
import importlib.util
import itertools
import os
import random

import pytest

_HERE = os.path.dirname(os.path.abspath(__file__))


def _real_posggym():
    try:
        spec = importlib.util.find_spec("posggym")
    except (ImportError, ValueError):
        return False
    if spec is None or not spec.origin:
        return False
    return os.path.join("oracle", "shims") not in spec.origin


pytestmark = pytest.mark.skipif(not _real_posggym(), reason="real posggym is not installed")

ENV_KWARGS = {   # baseline_exps/env_data/<env>/env_kwargs.yaml of the reference
    "Driving-v1": dict(grid="14x14RoundAbout", num_agents=2, obs_dim=(3, 1, 1)),
    "PredatorPrey-v0": dict(grid="10x10", num_predators=2, num_prey=3, cooperative=True, prey_strength=2, obs_dim=2),
    "LevelBasedForaging-v3": dict(num_agents=2, max_agent_level=3, size=10, max_food=8, sight=2, force_coop=False,
                                  static_layout=True, observation_mode="tuple"),
}


def _plain(x):
    """Nested tuples / numpy scalars -> plain nested tuples of ints (observations compare by value)."""
    if isinstance(x, (tuple, list)):
        return tuple(_plain(v) for v in x)
    try:
        return int(x)
    except (TypeError, ValueError):
        return x


def _outputs(ts, agents):
    return (tuple(_plain(ts.observations[i]) for i in agents), tuple(float(ts.rewards[i]) for i in agents),
            tuple(bool(ts.terminations[i]) for i in agents), bool(ts.all_done))


def _oracle_and_candidates(env_id):
    if env_id == "Driving-v1":
        from oracle.driving import DrivingModel

        m = DrivingModel()
        return m, (lambda s: [0, 1]), self_initial_driving
    if env_id == "PredatorPrey-v0":
        from oracle.predator_prey import PredatorPreyModel

        m = PredatorPreyModel()
        # ten words that reach every pick index floor(word * L / 2**32) of candidate lists of length L = 1..5
        fracs = [(0, 1), (1, 5), (1, 4), (1, 3), (2, 5), (1, 2), (3, 5), (2, 3), (3, 4), (4, 5)]
        words = [min((1 << 32) - 1, ((a << 32) // b) + 1) for a, b in fracs]
        return m, (lambda s: itertools.product(words, repeat=3)), self_initial_pp
    from oracle.lbf import LevelBasedForagingModel

    return LevelBasedForagingModel(), (lambda s: [None]), self_initial_lbf


def self_initial_driving(m, real_state):
    from oracle.driving import VehicleState

    out = []
    for v in real_state:
        coord, dest = tuple(v.coord), tuple(v.dest_coord)
        d0 = m.grid.get_shortest_path_distance(coord, dest)
        out.append(VehicleState(coord, int(v.facing_dir), int(v.speed), dest, int(v.dest_reached), int(v.crashed),
                                int(v.min_dest_dist), d0))
    return tuple(out)


def self_initial_pp(m, real_state):
    from oracle.predator_prey import PPState

    return PPState(tuple(tuple(c) for c in real_state.predator_coords), tuple(tuple(c) for c in real_state.prey_coords),
                   tuple(int(c) for c in real_state.prey_caught))


def self_initial_lbf(m, real_state, initial=False):
    from oracle.lbf import Food, LBFState, Player

    players = tuple(Player(int(p.idx), tuple(p.coord), int(p.level)) for p in real_state.players)
    food = tuple(Food(tuple(f.coord), int(f.level)) for f in real_state.food)
    assert [f.coord for f in food] == list(m.food_coords), "static layout differs from the oracle's guess"
    if initial:
        assert [p.coord for p in players] == list(m.player_coords), "player start cells differ from the oracle's guess"
    return LBFState(players, food, int(real_state.food_spawned))


@pytest.mark.parametrize("env_id", list(ENV_KWARGS))
def test_oracle_step_matches_real_posggym(env_id):
    import posggym

    real = posggym.make(env_id, **ENV_KWARGS[env_id]).model
    oracle, candidates, convert = _oracle_and_candidates(env_id)
    agents = tuple(real.possible_agents)
    assert agents == tuple(oracle.possible_agents)
    rng = random.Random(0)
    compared = 0
    for episode in range(60 if env_id == "PredatorPrey-v0" else 200):
        real.seed(episode)
        rs = real.sample_initial_state()
        os_ = convert(oracle, rs, True) if env_id == "LevelBasedForaging-v3" else convert(oracle, rs)
        ro = real.sample_initial_obs(rs)
        oo = oracle.sample_initial_obs(os_)
        assert tuple(_plain(ro[i]) for i in agents) == tuple(_plain(oo[i]) for i in agents)
        for t in range(50):
            acts = {i: rng.randrange(real.action_spaces[i].n) for i in agents}
            rts = real.step(rs, acts)
            want = _outputs(rts, agents)
            match = None
            for noise in candidates(os_):
                ots = oracle.step_with_noise(os_, acts, list(noise) if isinstance(noise, tuple) else noise)
                if _outputs(ots, agents) == want and (env_id == "Driving-v1" or ots.state == convert(oracle, rts.state)):
                    match = ots
                    break
            assert match is not None, (env_id, episode, t, "no oracle outcome equals the upstream step")
            rs, os_ = rts.state, match.state
            compared += 1
            if rts.all_done:
                break
    assert compared > 1500
