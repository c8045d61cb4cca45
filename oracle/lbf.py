"""LevelBasedForaging-v3 grid-world model (oracle; test infrastructure only).

Restates the published algorithm of ``posggym/envs/grid_world/level_based_foraging.py``
(posggym's port of Christianos et al.'s lb-foraging; un-vendored dependency
posggym>=0.5.0, /root/reference/pyproject.toml:36) for the reference's experiment kwargs
``num_agents=2, max_agent_level=3, size=10, max_food=8, sight=2, force_coop=False,
static_layout=True, observation_mode="tuple"``
(baseline_exps/env_data/LevelBasedForaging-v3/env_kwargs.yaml:1-10).

PARITY: unpinned against the upstream source (absent from the container).  What the
in-tree fixtures pin (tests/test_oracle_fixtures.py):

* the tuple observation = ``players x (x, y, level)`` (self first; coordinates relative to
  the (2*sight+1)^2 window, so self is always (sight, sight); an agent out of sight is
  (-1, -1, 0)) followed by ``max_food x (x, y, level)`` with the VISIBLE foods packed into
  the first slots; spaces Discrete(size+1, start=-1) / Discrete(max_agent_level+1) for
  players and Discrete(size+2, start=-1) / Discrete(2*max_agent_level+1) for foods, i.e. a
  flattened width of 2*26 + 8*31 = 300 (br_models/*.pt ``trunk.0.weight``); which one-hot
  inputs ever fire (Adam ``exp_avg_sq`` masks): self x/y only at value sight+1, other
  agent / food coordinates only in [0, 2*sight+1], food levels 0..6, at most six foods
  visible at once (slots 6 and 7 never filled);
* rewards are ``level * food_level / (sum of loading levels * total food level spawned)``:
  episode returns are small-denominator rationals in [0, 1] (``br_results.csv``);
* step limit 50.

Static layout.  Nothing in the reference tree states it; two fixtures select it:

* the ten shipped best-response checkpoints (``env_data/LevelBasedForaging-v3/br_models``, trained on the real
  environment) run on this model (tests/test_oracle_policies.py): with the foods on the LAST eight cells of the
  sorted lattice ``product(range(1, size-1, 2), repeat=2)`` -- x in {5, 7}, y in {1, 3, 5, 7} -- two of them in
  self-play collect 0.2 - 0.45 each from whatever border cells they start on -- they walk to that half of the
  field (the reference's ``br_results.csv``: 0.41 with its own partner population); with the first eight cells
  (x in {1, 3}: this file's first version) the agent that starts on the far side collects 0.00 in every episode,
  and with every other two-column / two-row / four-by-two / checkerboard placement tried the pair collects less
  than half as much; swapping the observation's (x, y) order gives ~0 everywhere.  The same runs confirm the observation layout, the action semantics (NORTH = y - 1) and
  the load rule.  The mask above (six foods visible at once) holds for this placement too;
* uniform-random play (``div_results.csv``: 0.0597 +- 0.0088 / 0.0158 +- 0.0041 -- agent 0 collects four times as
  much) puts agent 0 next to the foods, on the right border: see ``static_player_coords``.

Residual inconsistency (LBF stays the least certain of the three oracles): the shipped critics assign their highest
initial value (0.32 against 0.11 here) to a first observation with NOTHING in sight, and the policies' default motion
on such observations is EAST -- as if agent 0 started out of sight of the foods, west of them -- while the
uniform-random returns (and the first point of the training curves, 0.05) put agent 0 two cells from a food.  No
placement in the families tried (two adjacent columns / rows at every offset, 4 x 2 blocks, checkerboards; every
border start) satisfies both, so something else in this restatement -- the level sampling or the reward
normalisation -- may still differ from upstream (a diamond-shaped field of view was tried and is worse).  A second
family fits the uniform-random returns as well as this placement does: agent 0 in the top-left corner (0, 0) with the
nearest food two moves away (e.g. x in {1, 3}, y in {2, 4, 6, 8}, agent 1 at (9, 9)); there the trained agent 0
collects 0.29 against a random partner (0.20 here) and its critic starts at 0.26, but a second copy of the policy
playing agent 1 -- a role it was never trained in -- never finds the food.  The fixtures do not separate the two;
this placement is kept because both roles work on it.  tools/policy_probe.py reproduces the searches (``lbf``,
``lbf-random``, ``lbf-agent0``).

Still guesses: which of the two right-hand corners agent 0 starts in, and the level sampling (player levels uniform
on 1..max_agent_level, food levels uniform on 1..sum of player levels, as in lb-foraging).  The order in which
simultaneous LOADs are resolved follows lb-foraging (player index order, first adjacent food in N, S, W, E order).

Noise definition: ``step`` is deterministic.  ``sample_initial_state`` takes
``num_agents + max_food`` picks: player level - 1, then food level - 1.
"""
from __future__ import annotations

import itertools
import random
from typing import Dict, List, NamedTuple, Optional, Sequence, Tuple

Coord = Tuple[int, int]
NONE, NORTH, SOUTH, WEST, EAST, LOAD = 0, 1, 2, 3, 4, 5
ACTIONS = [NONE, NORTH, SOUTH, WEST, EAST, LOAD]
MOVE_DELTA = {NORTH: (0, -1), SOUTH: (0, 1), WEST: (-1, 0), EAST: (1, 0)}
ADJ_ORDER = ((0, -1), (0, 1), (-1, 0), (1, 0))   # N, S, W, E


class Player(NamedTuple):
    idx: int
    coord: Coord
    level: int


class Food(NamedTuple):
    coord: Coord
    level: int   # 0 = collected


class LBFState(NamedTuple):
    players: Tuple[Player, ...]
    food: Tuple[Food, ...]
    food_spawned: int


LBFObs = Tuple[int, ...]


def static_food_coords(size: int, max_food: int) -> List[Coord]:
    """The LAST ``max_food`` cells of the sorted spacing-2 interior lattice (size 10, 8 foods: x in {5, 7},
    y in {1, 3, 5, 7}) -- see the module docstring for what selects it."""
    cells = list(itertools.product(range(1, size - 1, 2), repeat=2))
    assert max_food <= len(cells)
    return cells[len(cells) - max_food:]


def static_player_coords(size: int, num_agents: int) -> List[Coord]:
    """Evenly spaced along the border, starting at the top-right corner and walking down the right side, along
    the bottom, up the left side and back along the top (two agents: (size-1, 0) and (0, size-1)).  With the foods
    on the right half this start reproduces the uniform-random returns of div_results.csv (0.060 / 0.016: agent
    0 collects more): 0.062 / 0.012; the mirrored pair (size-1, size-1) / (0, 0) fits equally (0.051 / 0.016),
    agent 0 on the left does not (0.013 / 0.060)."""
    n = size
    border = ([(n - 1, y) for y in range(n)] + [(x, n - 1) for x in range(n - 2, -1, -1)]
              + [(0, y) for y in range(n - 2, -1, -1)] + [(x, 0) for x in range(1, n - 1)])
    step = len(border) / num_agents
    return [border[int(i * step)] for i in range(num_agents)]


class LevelBasedForagingModel:
    is_symmetric = True

    def __init__(self, num_agents: int = 2, max_agent_level: int = 3, size: int = 10, max_food: int = 8,
                 sight: int = 2, force_coop: bool = False, static_layout: bool = True,
                 normalize_reward: bool = True, observation_mode: str = "tuple"):
        assert static_layout, "only the static layout of the reference's experiments is restated"
        assert observation_mode == "tuple"
        self.size, self.max_food, self.sight = int(size), int(max_food), int(sight)
        self.max_agent_level, self.force_coop = int(max_agent_level), bool(force_coop)
        self._normalize_reward = bool(normalize_reward)
        self.possible_agents = tuple(str(i) for i in range(num_agents))
        self.food_coords = static_food_coords(self.size, self.max_food)
        self.player_coords = static_player_coords(self.size, num_agents)
        self.spec = None
        self._rng: Optional[random.Random] = None

        from gymnasium import spaces

        n = num_agents
        max_food_level = self.max_agent_level * min(n, 4)
        comps = []
        for _ in range(n):
            comps += [spaces.Discrete(self.size + 1, start=-1), spaces.Discrete(self.size + 1, start=-1),
                      spaces.Discrete(self.max_agent_level + 1)]
        for _ in range(self.max_food):
            comps += [spaces.Discrete(self.size + 2, start=-1), spaces.Discrete(self.size + 2, start=-1),
                      spaces.Discrete(max_food_level + 1)]
        self.action_spaces = {i: spaces.Discrete(len(ACTIONS)) for i in self.possible_agents}
        self.observation_spaces = {i: spaces.Tuple(tuple(comps)) for i in self.possible_agents}

    # ---- POSGModel plumbing -----------------------------------------------------------
    @property
    def rng(self) -> random.Random:
        if self._rng is None:
            self._rng = random.Random()
        return self._rng

    def seed(self, seed: Optional[int] = None):
        self._rng = random.Random(seed)
        for k, sp in enumerate(self.action_spaces.values()):
            sp.seed(None if seed is None else seed + 1 + k)

    def get_agents(self, state) -> List[str]:
        return list(self.possible_agents)

    # ---- initial state ------------------------------------------------------------------
    def initial_noise_bounds(self, picks_so_far: Sequence[int]) -> int:
        n = len(self.possible_agents)
        if len(picks_so_far) < n:
            return self.max_agent_level
        return sum(p + 1 for p in picks_so_far[:n]) if not self.force_coop else 1

    def initial_state_from_noise(self, picks: Sequence[int]) -> LBFState:
        n = len(self.possible_agents)
        players = tuple(Player(i, self.player_coords[i], picks[i] + 1) for i in range(n))
        max_level = sum(p.level for p in players)
        foods = []
        for k in range(self.max_food):
            level = max_level if self.force_coop else picks[n + k] + 1
            foods.append(Food(self.food_coords[k], level))
        return LBFState(players, tuple(foods), sum(f.level for f in foods))

    def sample_initial_state(self) -> LBFState:
        picks: List[int] = []
        for _ in range(len(self.possible_agents) + self.max_food):
            picks.append(self.rng.randrange(self.initial_noise_bounds(picks)))
        return self.initial_state_from_noise(picks)

    def sample_initial_obs(self, state: LBFState) -> Dict[str, LBFObs]:
        return self._get_obs(state)

    def sample_agent_initial_state(self, agent_id: str, obs: LBFObs) -> LBFState:
        raise NotImplementedError

    # ---- step -----------------------------------------------------------------------------
    def step(self, state: LBFState, actions: Dict[str, int]):
        return self.step_with_noise(state, actions, None)

    def step_with_noise(self, state: LBFState, actions: Dict[str, int], noise=None):
        from posggym.model import JointTimestep, Outcome

        assert all(a in ACTIONS for a in actions.values())
        next_state, rewards = self._get_next_state(state, actions)
        obs = self._get_obs(next_state)
        all_done = all(f.level == 0 for f in next_state.food)
        terminated = {i: all_done for i in self.possible_agents}
        truncated = {i: False for i in self.possible_agents}
        infos = {i: {"outcome": Outcome.NA} for i in self.possible_agents}
        return JointTimestep(next_state, obs, rewards, terminated, truncated, all_done, infos)

    def _food_at(self, foods: Sequence[Food], c: Coord) -> int:
        for k, f in enumerate(foods):
            if f.level > 0 and f.coord == c:
                return k
        return -1

    def _adjacent_food(self, foods: Sequence[Food], c: Coord) -> int:
        for dx, dy in ADJ_ORDER:
            k = self._food_at(foods, (c[0] + dx, c[1] + dy))
            if k >= 0:
                return k
        return -1

    def _get_next_state(self, state: LBFState, actions: Dict[str, int]):
        n = len(self.possible_agents)
        foods = list(state.food)
        # 1. validity: moves need an in-bounds, food-free target; LOAD needs an adjacent food
        targets: List[Coord] = []
        loading: List[int] = []
        for p in state.players:
            a = actions[str(p.idx)]
            target = p.coord
            if a in MOVE_DELTA:
                dx, dy = MOVE_DELTA[a]
                c = (p.coord[0] + dx, p.coord[1] + dy)
                if 0 <= c[0] < self.size and 0 <= c[1] < self.size and self._food_at(foods, c) < 0:
                    target = c
            elif a == LOAD and self._adjacent_food(foods, p.coord) >= 0:
                loading.append(p.idx)
            targets.append(target)
        # 2. players whose target is claimed by more than one player stay where they are
        coords = []
        for i, p in enumerate(state.players):
            clash = sum(1 for t in targets if t == targets[i]) > 1
            coords.append(p.coord if clash else targets[i])
        players = [Player(p.idx, coords[i], p.level) for i, p in enumerate(state.players)]
        # 3. loading, in player index order
        rewards = {i: 0.0 for i in self.possible_agents}
        pending = list(loading)
        while pending:
            i = pending.pop(0)
            k = self._adjacent_food(foods, players[i].coord)
            if k < 0:
                continue
            fc = foods[k].coord
            group = [i] + [j for j in pending
                           if abs(players[j].coord[0] - fc[0]) + abs(players[j].coord[1] - fc[1]) == 1]
            pending = [j for j in pending if j not in group]
            level_sum = sum(players[j].level for j in group)
            if level_sum < foods[k].level:
                continue
            for j in group:
                r = float(players[j].level * foods[k].level)
                if self._normalize_reward:
                    r = r / float(level_sum * state.food_spawned)
                rewards[str(j)] += r
            foods[k] = Food(fc, 0)
        assert n == len(players)
        return LBFState(tuple(players), tuple(foods), state.food_spawned), rewards

    # ---- observations -------------------------------------------------------------------------
    def _get_obs(self, state: LBFState) -> Dict[str, LBFObs]:
        return {i: self._get_tuple_obs(int(i), state) for i in self.possible_agents}

    def _get_tuple_obs(self, idx: int, state: LBFState) -> LBFObs:
        me = state.players[idx]
        s = self.sight
        out: List[int] = [s, s, me.level]
        for p in state.players:
            if p.idx == idx:
                continue
            dx, dy = p.coord[0] - me.coord[0], p.coord[1] - me.coord[1]
            if abs(dx) <= s and abs(dy) <= s:
                out += [dx + s, dy + s, p.level]
            else:
                out += [-1, -1, 0]
        seen = 0
        for f in state.food:
            dx, dy = f.coord[0] - me.coord[0], f.coord[1] - me.coord[1]
            if f.level > 0 and abs(dx) <= s and abs(dy) <= s:
                out += [dx + s, dy + s, f.level]
                seen += 1
        out += [-1, -1, 0] * (self.max_food - seen)
        return tuple(out)
