"""LevelBasedForaging-v3 on the GPU: batched model + scalar ``POSGModel`` facade (call sites
planning/mcts.py:181,191-192,333,418; planning/belief.py:165; the PPO vec-env of
ppo/config.py:223-232).  Packed layouts: include/posg_b200.h, csrc/lbf.cuh.  Env kwargs of
the reference's experiments: baseline_exps/env_data/LevelBasedForaging-v3/env_kwargs.yaml."""
from __future__ import annotations

import ctypes as C
import itertools
from typing import List, NamedTuple, Tuple

import numpy as np
import torch

from posggym_baselines_b200 import _lib
from posggym_baselines_b200.envs.base import BatchedModel, ModelFacade, to_numpy
from posggym_baselines_b200.spaces import Discrete
from posggym_baselines_b200.spaces import Tuple as TupleSpace

NUM_ACTIONS = 6
MAX_EPISODE_STEPS = 50


class Player(NamedTuple):
    idx: int
    coord: Tuple[int, int]
    level: int


class Food(NamedTuple):
    coord: Tuple[int, int]
    level: int   # 0 = collected


class LBFState(NamedTuple):
    players: Tuple[Player, ...]
    food: Tuple[Food, ...]
    food_spawned: int


def static_food_coords(size: int, max_food: int) -> List[Tuple[int, int]]:
    """Static layout: the LAST ``max_food`` cells of the sorted spacing-2 interior lattice (size 10, 8 foods:
    x in {5, 7}, y in {1, 3, 5, 7}) -- the placement the reference's shipped best-response policies were trained
    on (oracle/lbf.py, tests/test_oracle_policies.py)."""
    cells = list(itertools.product(range(1, size - 1, 2), repeat=2))
    if max_food > len(cells):
        raise ValueError(f"max_food={max_food} does not fit a {size}x{size} static layout")
    return cells[len(cells) - max_food:]


def static_player_coords(size: int, num_agents: int) -> List[Tuple[int, int]]:
    """Static layout: players evenly spaced along the border, from the top-right corner down the right side."""
    n = size
    border = ([(n - 1, y) for y in range(n)] + [(x, n - 1) for x in range(n - 2, -1, -1)]
              + [(0, y) for y in range(n - 2, -1, -1)] + [(x, 0) for x in range(1, n - 1)])
    step = len(border) / num_agents
    return [border[int(i * step)] for i in range(num_agents)]


class BatchedLBF(BatchedModel):
    env_id = "LevelBasedForaging-v3"
    num_agents, state_words, obs_words, noise_words, num_picks = 2, 2, 2, 1, 10
    noise_is_dummy = True   # LevelBasedForaging's step takes no noise (one placeholder word in the C-ABI)

    def __init__(self, num_agents: int = 2, max_agent_level: int = 3, size: int = 10, max_food: int = 8,
                 sight: int = 2, force_coop: bool = False, static_layout: bool = True,
                 normalize_reward: bool = True, observation_mode: str = "tuple", device=None):
        if not static_layout:
            raise NotImplementedError("only static_layout=True (the reference's experiment kwargs) is built")
        if observation_mode != "tuple":
            raise NotImplementedError("only observation_mode='tuple' is built")
        self.size, self.max_food, self.sight = int(size), int(max_food), int(sight)
        self.max_agent_level, self.force_coop = int(max_agent_level), bool(force_coop)
        self.food_coords = static_food_coords(self.size, self.max_food)
        self.player_coords = static_player_coords(self.size, num_agents)
        self._init_device(device)
        food_xy = np.array([x | (y << 4) for x, y in self.food_coords], np.uint8)
        player_xy = np.array([x | (y << 4) for x, y in self.player_coords] + [0] * (4 - num_agents), np.uint8)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().posg_lbf_create(
                self.size, num_agents, self.max_agent_level, self.max_food, self.sight, int(self.force_coop),
                int(bool(normalize_reward)), food_xy.ctypes.data_as(C.c_void_p), player_xy.ctypes.data_as(C.c_void_p),
                C.byref(self._handle)))

    def initial_pick_bounds(self, picks_so_far) -> int:
        """Player levels (level - 1), then food levels (level - 1, at most the players' level sum)."""
        if len(picks_so_far) < self.num_agents:
            return self.max_agent_level
        return 1 if self.force_coop else sum(p + 1 for p in picks_so_far[: self.num_agents])

    # ---- codec ---------------------------------------------------------------------
    def pack_state(self, states) -> np.ndarray:
        out = np.zeros((len(states), 2), np.uint32)
        for i, s in enumerate(states):
            w0 = int(s.food_spawned) << 20
            for k, p in enumerate(s.players):
                w0 |= (p.coord[0] | (p.coord[1] << 4) | (p.level << 8)) << (10 * k)
            w1 = 0
            for k, f in enumerate(s.food):
                assert tuple(f.coord) == tuple(self.food_coords[k]), "static layout: food cells are fixed"
                w1 |= f.level << (3 * k)
            out[i] = (w0, w1)
        return out.view(np.int32)

    def unpack_state(self, packed) -> List[LBFState]:
        w = to_numpy(packed).view(np.uint32).astype(np.int64)
        out = []
        for w0, w1 in w:
            players = tuple(Player(k, (int((w0 >> (10 * k)) & 15), int((w0 >> (10 * k + 4)) & 15)),
                                   int((w0 >> (10 * k + 8)) & 3)) for k in range(2))
            food = tuple(Food(self.food_coords[k], int((w1 >> (3 * k)) & 7)) for k in range(self.max_food))
            out.append(LBFState(players, food, int((w0 >> 20) & 63)))
        return out

    @staticmethod
    def _triples(w0: int, w1: int) -> tuple:
        h = [w0 & 0xFFFFFFFF, w0 >> 32, w1 & 0xFFFFFFFF, w1 >> 32]
        out = [(h[0] & 7) - 1, ((h[0] >> 3) & 7) - 1, (h[0] >> 6) & 3,
               ((h[0] >> 8) & 7) - 1, ((h[0] >> 11) & 7) - 1, (h[0] >> 14) & 3]
        for slot in range(8):
            half = 0 if slot == 0 else (slot + 2) // 3
            off = 16 if slot == 0 else 9 * ((slot - 1) % 3)
            f = (h[half] >> off) & 511
            out += [(f & 7) - 1, ((f >> 3) & 7) - 1, (f >> 6) & 7]
        return tuple(out)

    @classmethod
    def pack_obs(cls, obs) -> np.ndarray:
        """Tuple observation (30 entries) -> the kernels' two packed words."""
        h = [0, 0, 0, 0]
        h[0] = (obs[0] + 1) | ((obs[1] + 1) << 3) | (obs[2] << 6) | ((obs[3] + 1) << 8) | ((obs[4] + 1) << 11) | (
            obs[5] << 14)
        for slot in range(8):
            x, y, lvl = obs[6 + 3 * slot: 9 + 3 * slot]
            half = 0 if slot == 0 else (slot + 2) // 3
            off = 16 if slot == 0 else 9 * ((slot - 1) % 3)
            h[half] |= ((x + 1) | ((y + 1) << 3) | (lvl << 6)) << off
        return np.array([h[0] | (h[1] << 32), h[2] | (h[3] << 32)], dtype=np.uint64)

    @classmethod
    def unpack_obs(cls, packed) -> List[tuple]:
        w = to_numpy(packed).view(np.uint64).astype(object)
        w = w.reshape(w.shape[0], w.shape[1], -1)
        return [tuple(cls._triples(int(o[0]), int(o[1])) for o in row) for row in w]


class LBFModel(ModelFacade):
    def __init__(self, num_agents: int = 2, max_agent_level: int = 3, size: int = 10, max_food: int = 8,
                 sight: int = 2, force_coop: bool = False, static_layout: bool = True,
                 normalize_reward: bool = True, observation_mode: str = "tuple", device=None):
        self._init_facade(BatchedLBF(num_agents, max_agent_level, size, max_food, sight, force_coop, static_layout,
                                     normalize_reward, observation_mode, device),
                          "LevelBasedForaging-v3", MAX_EPISODE_STEPS)
        self.action_spaces = {i: Discrete(NUM_ACTIONS) for i in self.possible_agents}
        comps = []
        for _ in range(num_agents):
            comps += [Discrete(size + 1, start=-1), Discrete(size + 1, start=-1), Discrete(max_agent_level + 1)]
        for _ in range(max_food):
            comps += [Discrete(size + 2, start=-1), Discrete(size + 2, start=-1),
                      Discrete(max_agent_level * min(num_agents, 4) + 1)]
        self.observation_spaces = {i: TupleSpace(tuple(comps)) for i in self.possible_agents}


LevelBasedForagingModel = LBFModel
