"""Driving-v1 grid-world model (oracle; test infrastructure only).

Restates the published algorithm of ``posggym/envs/grid_world/driving.py``
(``DrivingModel``; un-vendored dependency posggym>=0.5.0,
/root/reference/pyproject.toml:36) as consumed by the reference at
planning/mcts.py:181,191-192,333,418 and planning/belief.py:165, with every
random draw hoisted into an explicit noise argument.

PARITY: unpinned against the upstream source (absent from the container);
pinned against the in-tree fixtures (tests/test_oracle_fixtures.py):

* observation = ``(15 local cells, speed, (x, y), (dest x, dest y), dest_reached,
  crashed)`` -- flattened width 60+4+28+28+2+2 = 124 (br_models/*.pt
  ``trunk.0.weight``); component order, row-major 5x3 window with the ego cell at
  index 10 always VEHICLE, and the set of reachable destination one-hots all read
  off the Adam ``exp_avg_sq`` zero-columns of the shipped checkpoints;
* 14x14RoundAbout layout: the 16 (start, dest) shortest-path distances are
  {22:2, 23:4, 24:2, 25:4, 26:4} -- the histogram recovered from the float32
  reward residuals of 10 671 successful episodes in ``br_results.csv``;
* rewards: crash -1, progress ``0.5 * cells / D0`` (D0 = start->dest distance),
  destination 0.5 (so every success totals 1): ``float32(0.5 + k/(2*D0))``
  reproduces all ten distinct residuals of the successful-episode returns;
* single-vehicle dynamics: the minimum time-to-destination over the 16 (start,
  dest) pairs, grouped by D0, is {22:15, 23:15, 24:16, 25:16, 26:17} steps --
  exactly the minimum episode lengths of ``br_results.csv`` per D0 class (this
  also pins INIT_DIR to NORTH or its transpose-equivalent WEST);
* uniform-random play (``div_results.csv``): a lone random vehicle earns 0.104
  +- 0.005 progress reward per episode (oracle 0.104); random-vs-random crashes
  24/100 episodes with mean return -0.14 +- 0.045 (oracle 19.5 %, -0.10).  These
  statistics selected "a reversing vehicle keeps its facing" over "turn actions
  rotate a reversing vehicle" (16 % crashes, 0.116 progress).

* the shipped best-response checkpoints (``env_data/Driving-v1/br_models``, trained on the real environment)
  drive on this model: in self-play 95 % of the episodes reach the destination (``br_results.csv``: 88 % with the
  reference's partner population; uniform-random play: none) -- tests/test_oracle_policies.py.  The same check found
  defects in the first versions of the other two oracles; it found none here.

Noise definition (the "identical noise" of the bit-exact parity tests):

* ``step``: one integer ``order_idx`` in ``[0, n_agents!)`` -- the lexicographic
  index of the permutation in which vehicles are processed (upstream shuffles
  ``list(range(n))`` with ``model.rng``).  For two agents 0 -> (0, 1), 1 -> (1, 0).
* ``sample_initial_state``: per agent a pair ``(start_pick, dest_pick)`` indexing
  the still-available start / destination cells in grid parse order (row-major).
"""
from __future__ import annotations

import itertools
import math
import random
from typing import Dict, List, NamedTuple, Optional, Sequence, Set, Tuple

from oracle.grid import Coord, Direction, Grid

# --- constants --------------------------------------------------------------
REVERSE, STOPPED, FORWARD_SLOW, FORWARD_FAST = 0, 1, 2, 3
NUM_SPEEDS = 4
DO_NOTHING, ACCELERATE, DECELERATE, TURN_RIGHT, TURN_LEFT = 0, 1, 2, 3, 4
ACTIONS = [DO_NOTHING, ACCELERATE, DECELERATE, TURN_RIGHT, TURN_LEFT]
VEHICLE, WALL, EMPTY, DESTINATION = 0, 1, 2, 3
CELL_OBS = [VEHICLE, WALL, EMPTY, DESTINATION]
NO_COLLISION, OBSTACLE_COLLISION, VEHICLE_COLLISION = 0, 1, 2

INIT_DIR = Direction.NORTH
INIT_SPEED = STOPPED


class VehicleState(NamedTuple):
    coord: Coord
    facing_dir: int
    speed: int
    dest_coord: Coord
    dest_reached: int
    crashed: int
    min_dest_dist: int
    #: start -> destination shortest path distance; normaliser of the progress reward
    init_dest_dist: int


DState = Tuple[VehicleState, ...]
# obs = (local cells, speed, coord, dest coord, dest reached, crashed)
DObs = Tuple[Tuple[int, ...], int, Coord, Coord, int, int]


class DrivingGrid(Grid):
    """Grid with per-agent start and destination cells and cached shortest paths."""

    def __init__(self, width, height, block_coords, start_coords, dest_coords):
        super().__init__(width, height, block_coords)
        assert len(start_coords) == len(dest_coords)
        #: per agent index: cells in parse (row-major) order
        self.start_coords: List[List[Coord]] = [list(s) for s in start_coords]
        self.dest_coords: List[List[Coord]] = [list(d) for d in dest_coords]
        all_dests = sorted({c for d in self.dest_coords for c in d}, key=lambda c: (c[1], c[0]))
        self.shortest_paths = self.get_all_shortest_paths(all_dests)

    @property
    def supported_num_agents(self) -> int:
        return len(self.start_coords)

    def get_shortest_path_distance(self, coord: Coord, dest: Coord) -> int:
        return int(self.shortest_paths[dest][coord])


def parse_grid_str(grid_str: str, supported_num_agents: int) -> DrivingGrid:
    """``.`` empty, ``#`` block, ``0-9`` start of that agent, ``+`` start of any
    agent, ``a-j`` destination of that agent, ``-`` destination of any agent."""
    rows = [r for r in grid_str.splitlines() if r]
    assert all(len(r) == len(rows[0]) for r in rows)
    blocks: Set[Coord] = set()
    shared_start, shared_dest = [], []
    agent_start: Dict[int, List[Coord]] = {}
    agent_dest: Dict[int, List[Coord]] = {}
    for y, row in enumerate(rows):
        for x, ch in enumerate(row):
            c = (x, y)
            if ch == "#":
                blocks.add(c)
            elif ch == "+":
                shared_start.append(c)
            elif ch == "-":
                shared_dest.append(c)
            elif ch.isdigit():
                agent_start.setdefault(int(ch), []).append(c)
            elif ch in "abcdefghij":
                agent_dest.setdefault(ord(ch) - ord("a"), []).append(c)
            else:
                assert ch == ".", ch
    order = lambda cs: sorted(cs, key=lambda c: (c[1], c[0]))  # noqa: E731
    starts = [order(shared_start + agent_start.get(i, [])) for i in range(supported_num_agents)]
    dests = [order(shared_dest + agent_dest.get(i, [])) for i in range(supported_num_agents)]
    return DrivingGrid(len(rows[0]), len(rows), blocks, starts, dests)


GRID_14x14_ROUNDABOUT = (
    "#-..........-#\n"
    "-#####..#####+\n"
    ".#####..#####.\n"
    ".#####..#####.\n"
    ".###......###.\n"
    ".###......###.\n"
    "......##......\n"
    "......##......\n"
    ".###......###.\n"
    ".###......###.\n"
    ".#####..#####.\n"
    ".#####..#####.\n"
    "-#####..#####+\n"
    "#+..........+#\n"
)

GRID_7x7_ROUNDABOUT = (
    "#-...-#\n"
    "-##.##+\n"
    ".#...#.\n"
    "...#...\n"
    ".#...#.\n"
    "-##.##+\n"
    "#+...+#\n"
)

GRID_6x6_INTERSECTION = (
    "##0b##\n"
    "##..##\n"
    "d....3\n"
    "2....c\n"
    "##..##\n"
    "##a1##\n"
)

SUPPORTED_GRIDS = {
    "14x14RoundAbout": (GRID_14x14_ROUNDABOUT, 4),
    "7x7RoundAbout": (GRID_7x7_ROUNDABOUT, 4),
    "6x6Intersection": (GRID_6x6_INTERSECTION, 4),
}


def perm_from_index(n: int, idx: int) -> Tuple[int, ...]:
    """Lexicographic ``idx``-th permutation of ``range(n)`` (factorial number system)."""
    items = list(range(n))
    out = []
    for k in range(n, 0, -1):
        f = math.factorial(k - 1)
        q, idx = divmod(idx, f)
        out.append(items.pop(q))
    return tuple(out)


class DrivingModel:
    """Restated ``posggym.envs.grid_world.driving.DrivingModel`` (Driving-v1)."""

    R_STEP_COST = 0.0
    R_CRASH_VEHICLE = -1.0
    R_DESTINATION_REACHED = 0.5
    R_PROGRESS = 0.5

    is_symmetric = True

    def __init__(
        self,
        grid: str = "14x14RoundAbout",
        num_agents: int = 2,
        obs_dim: Sequence[int] = (3, 1, 1),
        obstacle_collisions: bool = False,
    ):
        if isinstance(grid, str):
            grid_str, n_supported = SUPPORTED_GRIDS[grid]
            grid = parse_grid_str(grid_str, n_supported)
        assert 0 < num_agents <= grid.supported_num_agents
        assert len(obs_dim) == 3 and obs_dim[0] > 0 and obs_dim[1] >= 0 and obs_dim[2] >= 0
        self.grid: DrivingGrid = grid
        self.obs_front, self.obs_back, self.obs_side = (int(v) for v in obs_dim)
        self.obs_dim = (self.obs_front, self.obs_back, self.obs_side)
        self._obstacle_collisions = obstacle_collisions
        self.possible_agents = tuple(str(i) for i in range(num_agents))
        self.spec = None
        self._rng: Optional[random.Random] = None

        from gymnasium import spaces  # shim or real; installed by oracle.shims.install()

        self.action_spaces = {i: spaces.Discrete(len(ACTIONS)) for i in self.possible_agents}
        depth = self.obs_front + self.obs_back + 1
        width = 2 * self.obs_side + 1
        self.observation_spaces = {
            i: spaces.Tuple(
                (
                    spaces.Tuple(tuple(spaces.Discrete(len(CELL_OBS)) for _ in range(depth * width))),
                    spaces.Discrete(NUM_SPEEDS),
                    spaces.Tuple((spaces.Discrete(grid.width), spaces.Discrete(grid.height))),
                    spaces.Tuple((spaces.Discrete(grid.width), spaces.Discrete(grid.height))),
                    spaces.Discrete(2),
                    spaces.Discrete(2),
                )
            )
            for i in self.possible_agents
        }

    # ---- POSGModel plumbing ------------------------------------------------
    @property
    def rng(self) -> random.Random:
        if self._rng is None:
            self._rng = random.Random()
        return self._rng

    def seed(self, seed: Optional[int] = None):
        self._rng = random.Random(seed)
        for k, sp in enumerate(self.action_spaces.values()):
            sp.seed(None if seed is None else seed + 1 + k)

    def get_agents(self, state: DState) -> List[str]:
        return list(self.possible_agents)

    # ---- initial state -----------------------------------------------------
    def num_initial_noise(self) -> int:
        return 2 * len(self.possible_agents)

    def initial_state_from_noise(self, picks: Sequence[int]) -> DState:
        """``picks[2*i]`` / ``picks[2*i+1]`` index the still-available start / dest
        cells of agent ``i`` (parse order); each must be < the number available."""
        state = []
        chosen_start: Set[Coord] = set()
        chosen_dest: Set[Coord] = set()
        for i in range(len(self.possible_agents)):
            avail_start = [c for c in self.grid.start_coords[i] if c not in chosen_start]
            start = avail_start[picks[2 * i]]
            chosen_start.add(start)
            avail_dest = [c for c in self.grid.dest_coords[i] if c not in chosen_dest and c != start]
            dest = avail_dest[picks[2 * i + 1]]
            chosen_dest.add(dest)
            d0 = self.grid.get_shortest_path_distance(start, dest)
            state.append(VehicleState(start, int(INIT_DIR), INIT_SPEED, dest, 0, 0, d0, d0))
        return tuple(state)

    def initial_noise_bounds(self, picks_so_far: Sequence[int]) -> int:
        """Number of choices for the next pick given the earlier ones."""
        k = len(picks_so_far)
        i, is_dest = divmod(k, 2)
        chosen_start: Set[Coord] = set()
        chosen_dest: Set[Coord] = set()
        for j in range(i):
            a = [c for c in self.grid.start_coords[j] if c not in chosen_start]
            s = a[picks_so_far[2 * j]]
            chosen_start.add(s)
            d = [c for c in self.grid.dest_coords[j] if c not in chosen_dest and c != s]
            chosen_dest.add(d[picks_so_far[2 * j + 1]])
        avail_start = [c for c in self.grid.start_coords[i] if c not in chosen_start]
        if not is_dest:
            return len(avail_start)
        s = avail_start[picks_so_far[2 * i]]
        return len([c for c in self.grid.dest_coords[i] if c not in chosen_dest and c != s])

    def sample_initial_state(self) -> DState:
        picks: List[int] = []
        for _ in range(self.num_initial_noise()):
            picks.append(self.rng.randrange(self.initial_noise_bounds(picks)))
        return self.initial_state_from_noise(picks)

    def sample_initial_obs(self, state: DState) -> Dict[str, DObs]:
        return self._get_obs(state)

    def sample_agent_initial_state(self, agent_id: str, obs: DObs) -> DState:
        raise NotImplementedError

    # ---- step ----------------------------------------------------------------
    def num_step_orders(self) -> int:
        return math.factorial(len(self.possible_agents))

    def step(self, state: DState, actions: Dict[str, int]):
        return self.step_with_noise(state, actions, self.rng.randrange(self.num_step_orders()))

    def step_with_noise(self, state: DState, actions: Dict[str, int], order_idx: int):
        from posggym.model import JointTimestep, Outcome

        assert all(a in ACTIONS for a in actions.values())
        exec_order = perm_from_index(len(self.possible_agents), order_idx)
        next_state, collisions = self._get_next_state(state, actions, exec_order)
        obs = self._get_obs(next_state)
        rewards = self._get_rewards(state, next_state, collisions)
        terminated = {
            i: bool(next_state[int(i)].dest_reached or next_state[int(i)].crashed)
            for i in self.possible_agents
        }
        truncated = {i: False for i in self.possible_agents}
        all_done = all(terminated.values())
        infos = {}
        for i in self.possible_agents:
            vs = next_state[int(i)]
            outcome = Outcome.WIN if vs.dest_reached else (Outcome.LOSS if vs.crashed else Outcome.NA)
            infos[i] = {"outcome": outcome}
        return JointTimestep(next_state, obs, rewards, terminated, truncated, all_done, infos)

    def _get_next_state(self, state: DState, actions: Dict[str, int], exec_order: Sequence[int]):
        next_state = list(state)
        vehicle_coords = {vs.coord for vs in state}
        collisions = [NO_COLLISION] * len(state)
        for idx in exec_order:
            state_i = state[idx]
            if state_i.dest_reached or next_state[idx].crashed:
                # finished earlier, crashed earlier, or was hit earlier in this step
                continue
            action_i = actions[str(idx)]
            vehicle_coords.remove(state_i.coord)

            next_speed = self._get_next_speed(action_i, state_i.speed)
            move_dir = self._get_move_direction(action_i, next_speed, state_i.facing_dir)
            # "no turning while in reverse": a reversing vehicle keeps its facing
            next_dir = (
                state_i.facing_dir
                if next_speed == REVERSE
                else self._get_next_direction(action_i, state_i.facing_dir)
            )
            next_coord, collision, hit_coord = self._get_next_coord(
                state_i.coord, next_speed, move_dir, vehicle_coords
            )
            if collision == VEHICLE_COLLISION:
                for jdx in range(len(next_state)):
                    if next_state[jdx].coord == hit_coord:
                        collisions[jdx] = VEHICLE_COLLISION
                        next_state[jdx] = next_state[jdx]._replace(speed=STOPPED, crashed=1)
                        break

            crashed = collision == VEHICLE_COLLISION or (
                collision == OBSTACLE_COLLISION and self._obstacle_collisions
            )
            collisions[idx] = collision
            if crashed:
                next_speed = STOPPED

            min_dest_dist = min(
                state_i.min_dest_dist,
                self.grid.get_shortest_path_distance(next_coord, state_i.dest_coord),
            )
            next_state[idx] = VehicleState(
                coord=next_coord,
                facing_dir=next_dir,
                speed=next_speed,
                dest_coord=state_i.dest_coord,
                dest_reached=int(next_coord == state_i.dest_coord),
                crashed=int(crashed),
                min_dest_dist=min_dest_dist,
                init_dest_dist=state_i.init_dest_dist,
            )
            vehicle_coords.add(next_coord)
        return tuple(next_state), collisions

    @staticmethod
    def _get_next_direction(action: int, curr_dir: int) -> int:
        if action == TURN_RIGHT:
            return (curr_dir + 1) % 4
        if action == TURN_LEFT:
            return (curr_dir - 1) % 4
        return curr_dir

    def _get_move_direction(self, action: int, speed: int, curr_dir: int) -> int:
        if speed == REVERSE:
            # no turning while reversing: move opposite to the current facing
            return (curr_dir + 2) % 4
        return self._get_next_direction(action, curr_dir)

    @staticmethod
    def _get_next_speed(action: int, curr_speed: int) -> int:
        if action == DO_NOTHING:
            return curr_speed
        if action in (TURN_LEFT, TURN_RIGHT):
            return FORWARD_SLOW if curr_speed == FORWARD_FAST else curr_speed
        if action == ACCELERATE:
            return min(curr_speed + 1, FORWARD_FAST)
        return max(curr_speed - 1, REVERSE)

    def _get_next_coord(self, curr: Coord, speed: int, move_dir: int, vehicle_coords: Set[Coord]):
        next_coord = curr
        collision = NO_COLLISION
        hit = None
        for _ in range(abs(speed - STOPPED)):
            next_coord = self.grid.get_next_coord(curr, move_dir, ignore_blocks=False)
            if next_coord == curr:
                collision = OBSTACLE_COLLISION
                break
            if next_coord in vehicle_coords:
                collision = VEHICLE_COLLISION
                hit = next_coord
                next_coord = curr
                break
            curr = next_coord
        return next_coord, collision, hit

    # ---- rewards ---------------------------------------------------------------
    def _get_rewards(self, state: DState, next_state: DState, collisions) -> Dict[str, float]:
        rewards = {}
        for i in self.possible_agents:
            idx = int(i)
            if state[idx].crashed or state[idx].dest_reached:
                rewards[i] = 0.0
                continue
            if collisions[idx] == VEHICLE_COLLISION or (
                self._obstacle_collisions and collisions[idx] == OBSTACLE_COLLISION
            ):
                r = self.R_CRASH_VEHICLE
            elif next_state[idx].dest_reached:
                r = self.R_DESTINATION_REACHED
            else:
                r = self.R_STEP_COST
            progress = state[idx].min_dest_dist - next_state[idx].min_dest_dist
            r += self.R_PROGRESS * (max(0, progress) / max(1, state[idx].init_dest_dist))
            rewards[i] = r
        return rewards

    # ---- observations -------------------------------------------------------------
    def _get_obs(self, state: DState) -> Dict[str, DObs]:
        obs = {}
        for i in self.possible_agents:
            vs = state[int(i)]
            obs[i] = (
                self._get_local_cell_obs(int(i), state),
                vs.speed,
                vs.coord,
                vs.dest_coord,
                vs.dest_reached,
                vs.crashed,
            )
        return obs

    def _get_local_cell_obs(self, idx: int, state: DState) -> Tuple[int, ...]:
        depth = self.obs_front + self.obs_back + 1
        width = 2 * self.obs_side + 1
        vs = state[idx]
        vehicle_coords = {v.coord for v in state}
        cells = []
        for row, col in itertools.product(range(depth), range(width)):
            c = self._map_obs_to_grid_coord((col, row), vs.coord, vs.facing_dir)
            if c is None or c in self.grid.block_coords:
                cells.append(WALL)
            elif c in vehicle_coords:
                cells.append(VEHICLE)
            elif c == vs.dest_coord:
                cells.append(DESTINATION)
            else:
                cells.append(EMPTY)
        return tuple(cells)

    def _map_obs_to_grid_coord(self, obs_coord: Coord, agent_coord: Coord, facing_dir: int):
        agent_col, agent_row = agent_coord
        obs_col, obs_row = obs_coord
        if facing_dir == Direction.NORTH:
            row = agent_row + obs_row - self.obs_front
            col = agent_col + obs_col - self.obs_side
        elif facing_dir == Direction.EAST:
            row = agent_row + obs_col - self.obs_side
            col = agent_col - obs_row + self.obs_front
        elif facing_dir == Direction.SOUTH:
            row = agent_row - obs_row + self.obs_front
            col = agent_col - obs_col + self.obs_side
        else:
            row = agent_row - obs_col + self.obs_side
            col = agent_col + obs_row - self.obs_front
        if 0 <= row < self.grid.height and 0 <= col < self.grid.width:
            return (col, row)
        return None
