"""PredatorPrey-v0 grid-world model (oracle; test infrastructure only).

Restates the published algorithm of ``posggym/envs/grid_world/predator_prey.py``
(``PredatorPreyModel``; un-vendored dependency posggym>=0.5.0,
/root/reference/pyproject.toml:36) with every random draw hoisted into an explicit
noise argument.  Env kwargs of the reference's experiments:
``grid="10x10", num_predators=2, num_prey=3, cooperative=True, prey_strength=2,
obs_dim=2`` (baseline_exps/env_data/PredatorPrey-v0/env_kwargs.yaml:1-8).

PARITY: unpinned against the upstream source (absent from the container); pinned against
the in-tree fixtures (tests/test_oracle_fixtures.py):

* observation = 25 local cells, each EMPTY=0 / WALL=1 / PREDATOR=2 / PREY=3 -- flattened
  width 100 (br_models/*.pt ``trunk.0.weight``); the centre cell (index 12) is always
  PREDATOR and every other one-hot input occurs (Adam ``exp_avg_sq`` masks);
* cell ORDER and type codes, action semantics and the dynamics as a whole: the ten shipped best-response
  checkpoints (``env_data/PredatorPrey-v0/br_models``, trained on the real environment) are run on this model
  (tests/test_oracle_policies.py).  With the window in ROW-major order (index = row * 5 + col, row = y offset)
  two of them in self-play collect a mean return of 0.89 (the reference's ``br_results.csv``: 0.75 with its own
  partner population); with the window transposed -- this file's first version, column-major -- or with any other
  symmetry of the window or permutation of the type codes they collect 0.00 to 0.36;
* returns live on the lattice {0, 1/3, 2/3, 1} (``br_results.csv``): cooperative reward
  ``n_caught * (1 / num_prey)`` shared by all predators; episodes end when all prey are
  caught; step limit 50;
* uniform-random predators catch nothing: return 0.000 +- 0.000 over 100 episodes
  (``div_results.csv``).

Dynamics (as published): prey move first, one at a time in index order -- away from the
closest predator within ``obs_dim`` (Manhattan), else away from the closest other prey
within ``obs_dim``, else to a uniformly random free neighbouring cell or staying; "away"
= the free neighbour (or the current cell) with the largest distance, ties to the largest
``(dist, (x, y))`` tuple.  Predators then move simultaneously; a predator stays when its
target holds a prey or when another predator has the same target.  A prey is caught when
at least ``prey_strength`` predators are within Manhattan distance 1.

Noise definition: ``step`` takes one 32-bit word per prey; a prey uses its word for the
single uniform choice it makes that step (which of several equally close predators / prey
to flee from, or which random move): ``index = floor(word * len(candidates) / 2**32)``.
``sample_initial_state`` takes ``num_predators + num_prey`` picks without replacement
from the start-cell lists (upstream shuffles the lists and takes a prefix).
"""
from __future__ import annotations

import itertools
import random
from typing import Dict, List, NamedTuple, Optional, Sequence, Tuple

from oracle.grid import Coord, Direction, Grid

DO_NOTHING, UP, DOWN, LEFT, RIGHT = 0, 1, 2, 3, 4
ACTIONS = [DO_NOTHING, UP, DOWN, LEFT, RIGHT]
ACTION_TO_DIR = [None, Direction.NORTH, Direction.SOUTH, Direction.WEST, Direction.EAST]
EMPTY, WALL, PREDATOR, PREY = 0, 1, 2, 3
CELL_OBS = [EMPTY, WALL, PREDATOR, PREY]


class PPState(NamedTuple):
    predator_coords: Tuple[Coord, ...]
    prey_coords: Tuple[Coord, ...]
    prey_caught: Tuple[int, ...]


PPObs = Tuple[int, ...]


class PPGrid(Grid):
    def __init__(self, grid_size: int, block_coords=None, predator_start_coords=None, prey_start_coords=None):
        assert grid_size >= 3
        super().__init__(grid_size, grid_size, block_coords)
        self.size = grid_size
        if predator_start_coords is None:
            # corners and half-way along each side
            predator_start_coords = [
                c for c in itertools.product([0, grid_size // 2, grid_size - 1], repeat=2)
                if c[0] in (0, grid_size - 1) or c[1] in (0, grid_size - 1)
            ]
        self.predator_start_coords: List[Coord] = list(predator_start_coords)
        self.prey_start_coords: Optional[List[Coord]] = None if prey_start_coords is None else list(prey_start_coords)

    def get_unblocked_center_coords(self, num: int) -> List[Coord]:
        """At least ``num`` unblocked cells closest to the centre (whole rings at a time)."""
        cx, cy = self.width // 2, self.height // 2
        out: List[Coord] = []
        ring = 0
        while len(out) < num:
            for x in range(cx - ring, cx + ring + 1):
                for y in range(cy - ring, cy + ring + 1):
                    if max(abs(x - cx), abs(y - cy)) != ring:
                        continue
                    c = (x, y)
                    if self.coord_in_bounds(c) and c not in self.block_coords:
                        out.append(c)
            ring += 1
        return out


def get_grid(name: str) -> PPGrid:
    size = {"5x5": 5, "10x10": 10, "15x15": 15, "20x20": 20}[name]
    return PPGrid(size, None)


def pick_index(word: int, n: int) -> int:
    return (int(word) * n) >> 32


class PredatorPreyModel:
    R_MAX = 1.0
    is_symmetric = True

    def __init__(self, grid: str = "10x10", num_predators: int = 2, num_prey: int = 3,
                 cooperative: bool = True, prey_strength: Optional[int] = None, obs_dim: int = 2):
        if isinstance(grid, str):
            grid = get_grid(grid)
        if prey_strength is None:
            prey_strength = min(4, num_predators)
        assert 1 < num_predators <= 8 and num_prey > 0 and obs_dim > 0
        assert 0 < prey_strength <= min(4, num_predators)
        self.grid: PPGrid = grid
        self.obs_dim = int(obs_dim)
        self.num_prey = int(num_prey)
        self.cooperative = bool(cooperative)
        self.prey_strength = int(prey_strength)
        self._per_prey_reward = self.R_MAX / self.num_prey
        if self.grid.prey_start_coords is None:
            self.grid.prey_start_coords = self.grid.get_unblocked_center_coords(num_prey)
        self.possible_agents = tuple(str(i) for i in range(num_predators))
        self.spec = None
        self._rng: Optional[random.Random] = None

        from gymnasium import spaces

        n_cells = (2 * self.obs_dim + 1) ** 2
        self.action_spaces = {i: spaces.Discrete(len(ACTIONS)) for i in self.possible_agents}
        self.observation_spaces = {
            i: spaces.Tuple(tuple(spaces.Discrete(len(CELL_OBS)) for _ in range(n_cells)))
            for i in self.possible_agents
        }

    # ---- POSGModel plumbing ------------------------------------------------------
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

    # ---- initial state ------------------------------------------------------------
    def initial_noise_bounds(self) -> List[int]:
        np_, nq = len(self.grid.predator_start_coords), len(self.grid.prey_start_coords)
        return [np_ - k for k in range(len(self.possible_agents))] + [nq - k for k in range(self.num_prey)]

    def initial_state_from_noise(self, picks: Sequence[int]) -> PPState:
        pred_pool = list(self.grid.predator_start_coords)
        prey_pool = list(self.grid.prey_start_coords)
        n_pred = len(self.possible_agents)
        preds = [pred_pool.pop(picks[k]) for k in range(n_pred)]
        prey = [prey_pool.pop(picks[n_pred + k]) for k in range(self.num_prey)]
        return PPState(tuple(preds), tuple(prey), (0,) * self.num_prey)

    def sample_initial_state(self) -> PPState:
        return self.initial_state_from_noise([self.rng.randrange(b) for b in self.initial_noise_bounds()])

    def sample_initial_obs(self, state: PPState) -> Dict[str, PPObs]:
        return self._get_obs(state)

    def sample_agent_initial_state(self, agent_id: str, obs: PPObs) -> PPState:
        raise NotImplementedError

    # ---- step -----------------------------------------------------------------------
    def step(self, state: PPState, actions: Dict[str, int]):
        return self.step_with_noise(state, actions, [self.rng.getrandbits(32) for _ in range(self.num_prey)])

    def step_with_noise(self, state: PPState, actions: Dict[str, int], words: Sequence[int]):
        from posggym.model import JointTimestep, Outcome

        assert all(a in ACTIONS for a in actions.values())
        next_prey = self._get_next_prey_state(state, words)
        next_pred = self._get_next_predator_state(state, actions, next_prey)
        next_caught = self._get_next_prey_caught(state, next_prey, next_pred)
        next_state = PPState(next_pred, next_prey, next_caught)
        obs = self._get_obs(next_state)
        rewards = self._get_rewards(state, next_state)
        all_done = all(next_state.prey_caught)
        terminated = {i: all_done for i in self.possible_agents}
        truncated = {i: False for i in self.possible_agents}
        infos = {i: {"outcome": Outcome.WIN if all_done else Outcome.NA} for i in self.possible_agents}
        return JointTimestep(next_state, obs, rewards, terminated, truncated, all_done, infos)

    def _get_next_prey_state(self, state: PPState, words: Sequence[int]) -> Tuple[Coord, ...]:
        next_coords: List[Optional[Coord]] = [None] * self.num_prey
        occupied = set(state.predator_coords) | {
            c for i, c in enumerate(state.prey_coords) if not state.prey_caught[i]
        }
        # 1. flee from the closest predator in sight
        for i in range(self.num_prey):
            coord = state.prey_coords[i]
            if state.prey_caught[i]:
                next_coords[i] = coord
                continue
            nxt = self._move_away(coord, list(state.predator_coords), occupied, words[i])
            if nxt is not None:
                next_coords[i] = nxt
                occupied.remove(coord)
                occupied.add(nxt)
        # 2. keep away from the closest other prey in sight
        if self.num_prey - sum(state.prey_caught) > 1:
            for i in range(self.num_prey):
                if next_coords[i] is not None:
                    continue
                coord = state.prey_coords[i]
                others = [c for j, c in enumerate(state.prey_coords) if j != i and not state.prey_caught[j]]
                nxt = self._move_away(coord, others, occupied, words[i])
                if nxt is not None:
                    next_coords[i] = nxt
                    occupied.remove(coord)
                    occupied.add(nxt)
        # 3. nobody in sight: random move (free neighbours N, E, S, W, then staying)
        for i in range(self.num_prey):
            if next_coords[i] is not None:
                continue
            coord = state.prey_coords[i]
            cands = [c for c in self.grid.get_neighbours(coord) if c not in occupied] + [coord]
            nxt = cands[pick_index(words[i], len(cands))]
            next_coords[i] = nxt
            occupied.remove(coord)
            occupied.add(nxt)
        return tuple(next_coords)  # type: ignore[arg-type]

    def _move_away(self, coord: Coord, threats: List[Coord], occupied, word: int) -> Optional[Coord]:
        if not threats:
            return None
        dists = [Grid.manhattan_dist(coord, c) for c in threats]
        dmin = min(dists)
        if dmin > self.obs_dim:
            return None
        closest = [c for c, d in zip(threats, dists) if d == dmin]
        threat = closest[pick_index(word, len(closest))]
        options = [(Grid.manhattan_dist(coord, threat), coord)]
        for c in self.grid.get_neighbours(coord):
            if c in occupied:
                continue
            options.append((Grid.manhattan_dist(c, threat), c))
        options.sort()
        return options[-1][1]

    def _get_next_predator_state(self, state: PPState, actions: Dict[str, int], next_prey) -> Tuple[Coord, ...]:
        prey_cells = {c for i, c in enumerate(next_prey) if not state.prey_caught[i]}
        potential = []
        for i, coord in enumerate(state.predator_coords):
            a = actions[str(i)]
            nxt = coord if a == DO_NOTHING else self.grid.get_next_coord(coord, ACTION_TO_DIR[a], ignore_blocks=False)
            if nxt in prey_cells:
                nxt = coord
            potential.append(nxt)
        out = []
        for i, c in enumerate(potential):
            if any(j != i and potential[j] == c for j in range(len(potential))):
                c = state.predator_coords[i]
            out.append(c)
        return tuple(out)

    def _get_next_prey_caught(self, state: PPState, next_prey, next_pred) -> Tuple[int, ...]:
        caught = []
        for i in range(self.num_prey):
            if state.prey_caught[i]:
                caught.append(1)
                continue
            adjacent = sum(Grid.manhattan_dist(next_prey[i], c) <= 1 for c in next_pred)
            caught.append(int(adjacent >= self.prey_strength))
        return tuple(caught)

    def _get_rewards(self, state: PPState, next_state: PPState) -> Dict[str, float]:
        new_caught = [
            next_state.prey_coords[i] for i in range(self.num_prey)
            if not state.prey_caught[i] and next_state.prey_caught[i]
        ]
        if not new_caught:
            return {i: 0.0 for i in self.possible_agents}
        if self.cooperative:
            r = len(new_caught) * self._per_prey_reward
            return {i: r for i in self.possible_agents}
        rewards = {i: 0.0 for i in self.possible_agents}
        for prey in new_caught:
            involved = [i for i, c in zip(self.possible_agents, next_state.predator_coords)
                        if Grid.manhattan_dist(prey, c) <= 1]
            for i in involved:
                rewards[i] += self._per_prey_reward / len(involved)
        return rewards

    # ---- observations -----------------------------------------------------------------
    def _get_obs(self, state: PPState) -> Dict[str, PPObs]:
        return {i: self._get_local_cell_obs(int(i), state) for i in self.possible_agents}

    def _get_local_cell_obs(self, idx: int, state: PPState) -> PPObs:
        width = 2 * self.obs_dim + 1
        ax, ay = state.predator_coords[idx]
        preds = set(state.predator_coords)
        prey = {c for i, c in enumerate(state.prey_coords) if not state.prey_caught[i]}
        cells = []
        for row, col in itertools.product(range(width), repeat=2):     # row-major: index = row * width + col
            c = (ax + col - self.obs_dim, ay + row - self.obs_dim)
            if not self.grid.coord_in_bounds(c) or c in self.grid.block_coords:
                cells.append(WALL)
            elif c in preds:
                cells.append(PREDATOR)
            elif c in prey:
                cells.append(PREY)
            else:
                cells.append(EMPTY)
        return tuple(cells)
