"""Grid helpers (oracle; test infrastructure only).

Restates ``posggym/envs/grid_world/core.py`` (Direction, Coord, Grid) of the
un-vendored dependency posggym>=0.5.0 (/root/reference/pyproject.toml:36).
Coordinates are ``(x, y) = (col, row)`` with the origin in the top-left cell.
"""
from __future__ import annotations

import enum
from collections import deque
from typing import Dict, Iterable, List, Optional, Set, Tuple

Coord = Tuple[int, int]


class Direction(enum.IntEnum):
    NORTH = 0
    EAST = 1
    SOUTH = 2
    WEST = 3


#: (dx, dy) per Direction value
DIR_DELTA = ((0, -1), (1, 0), (0, 1), (-1, 0))


class Grid:
    """A rectangular grid with blocked cells."""

    def __init__(self, width: int, height: int, block_coords: Optional[Iterable[Coord]] = None):
        self.width = width
        self.height = height
        self.block_coords: Set[Coord] = set(block_coords) if block_coords else set()

    # ------------------------------------------------------------------
    @property
    def all_coords(self) -> List[Coord]:
        return [(x, y) for y in range(self.height) for x in range(self.width)]

    @property
    def unblocked_coords(self) -> List[Coord]:
        return [c for c in self.all_coords if c not in self.block_coords]

    def coord_in_bounds(self, coord: Coord) -> bool:
        return 0 <= coord[0] < self.width and 0 <= coord[1] < self.height

    @staticmethod
    def manhattan_dist(a: Coord, b: Coord) -> int:
        return abs(a[0] - b[0]) + abs(a[1] - b[1])

    # ------------------------------------------------------------------
    def get_next_coord(self, coord: Coord, move_dir: int, ignore_blocks: bool = False) -> Coord:
        """Next cell in ``move_dir``; the same cell when leaving the grid or
        (unless ``ignore_blocks``) entering a block."""
        x, y = coord
        if move_dir == Direction.NORTH:
            y = max(0, y - 1)
        elif move_dir == Direction.EAST:
            x = min(self.width - 1, x + 1)
        elif move_dir == Direction.SOUTH:
            y = min(self.height - 1, y + 1)
        elif move_dir == Direction.WEST:
            x = max(0, x - 1)
        new = (x, y)
        if not ignore_blocks and new in self.block_coords:
            return coord
        return new

    def get_neighbours(
        self, coord: Coord, ignore_blocks: bool = False, include_out_of_bounds: bool = False
    ) -> List[Coord]:
        """Neighbouring cells in the order N, E, S, W."""
        x, y = coord
        out: List[Coord] = []
        if y > 0 or include_out_of_bounds:
            out.append((x, y - 1))
        if x < self.width - 1 or include_out_of_bounds:
            out.append((x + 1, y))
        if y < self.height - 1 or include_out_of_bounds:
            out.append((x, y + 1))
        if x > 0 or include_out_of_bounds:
            out.append((x - 1, y))
        if ignore_blocks:
            return out
        return [c for c in out if c not in self.block_coords]

    # ------------------------------------------------------------------
    def dijkstra(self, origin: Coord) -> Dict[Coord, int]:
        """Unit-cost shortest path distance from ``origin`` to every reachable cell."""
        dist = {origin: 0}
        q = deque([origin])
        while q:
            c = q.popleft()
            for n in self.get_neighbours(c, ignore_blocks=False):
                if n not in dist:
                    dist[n] = dist[c] + 1
                    q.append(n)
        return dist

    def get_all_shortest_paths(self, origins: Iterable[Coord]) -> Dict[Coord, Dict[Coord, int]]:
        return {o: self.dijkstra(o) for o in origins}
