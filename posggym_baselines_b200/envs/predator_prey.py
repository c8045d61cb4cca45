"""PredatorPrey-v0 on the GPU: batched model + scalar ``POSGModel`` facade (call sites
planning/mcts.py:181,191-192,333,418; planning/belief.py:165).  Packed layouts:
include/posg_b200.h."""
from __future__ import annotations

import ctypes as C
from typing import List, NamedTuple, Optional, Tuple

import numpy as np
import torch

from posggym_baselines_b200 import _lib
from posggym_baselines_b200.envs.base import BatchedModel, ModelFacade, to_numpy
from posggym_baselines_b200.spaces import Discrete
from posggym_baselines_b200.spaces import Tuple as TupleSpace

NUM_ACTIONS = 5
MAX_EPISODE_STEPS = 50
GRID_SIZES = {"5x5": 5, "10x10": 10, "15x15": 15}


class PPState(NamedTuple):
    predator_coords: Tuple[Tuple[int, int], ...]
    prey_coords: Tuple[Tuple[int, int], ...]
    prey_caught: Tuple[int, ...]


class BatchedPredatorPrey(BatchedModel):
    env_id = "PredatorPrey-v0"
    num_agents, state_words, obs_words, noise_words, num_picks = 2, 2, 1, 3, 5

    def __init__(self, grid: str = "10x10", num_predators: int = 2, num_prey: int = 3, cooperative: bool = True,
                 prey_strength: Optional[int] = None, obs_dim: int = 2, device=None):
        self.size = GRID_SIZES[grid] if isinstance(grid, str) else int(grid)
        self.num_prey = num_prey
        self.obs_dim = obs_dim
        if prey_strength is None:
            prey_strength = min(4, num_predators)
        self._init_device(device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().posg_pp_create(self.size, None, num_predators, num_prey, int(cooperative),
                                                 prey_strength, obs_dim, C.byref(self._handle)))

    def initial_pick_bounds(self, picks_so_far) -> int:
        """Predators then prey, drawn without replacement from their start-cell lists (8 border cells; whole
        rings around the centre until there are at least ``num_prey`` cells)."""
        k = len(picks_so_far)
        n_pred_start = 8
        n_prey_start, ring = 0, 0
        while n_prey_start < self.num_prey:
            n_prey_start += 1 if ring == 0 else 8 * ring
            ring += 1
        return n_pred_start - k if k < self.num_agents else n_prey_start - (k - self.num_agents)

    # ---- codec ---------------------------------------------------------------------
    def pack_state(self, states) -> np.ndarray:
        out = np.zeros((len(states), 2), np.uint32)
        for i, s in enumerate(states):
            p, q, c = s.predator_coords, s.prey_coords, s.prey_caught
            out[i, 0] = (p[0][0] | (p[0][1] << 4)) | ((p[1][0] | (p[1][1] << 4)) << 8) | (
                (c[0] | (c[1] << 1) | (c[2] << 2)) << 16)
            out[i, 1] = (q[0][0] | (q[0][1] << 4)) | ((q[1][0] | (q[1][1] << 4)) << 8) | (
                (q[2][0] | (q[2][1] << 4)) << 16)
        return out.view(np.int32)

    def unpack_state(self, packed) -> List[PPState]:
        w = to_numpy(packed).view(np.uint32).astype(np.int64)
        xy = lambda v: (int(v & 15), int((v >> 4) & 15))  # noqa: E731
        return [PPState((xy(a), xy(a >> 8)), (xy(b), xy(b >> 8), xy(b >> 16)),
                        (int((a >> 16) & 1), int((a >> 17) & 1), int((a >> 18) & 1))) for a, b in w]

    @staticmethod
    def pack_obs(obs) -> np.ndarray:
        o = 0
        for c, v in enumerate(obs):
            o |= int(v) << (2 * c)
        return np.array([o], dtype=np.uint64)

    @staticmethod
    def unpack_obs(packed) -> List[tuple]:
        w = to_numpy(packed).astype(np.int64)
        w = w.reshape(w.shape[0], w.shape[1], -1)
        return [tuple(tuple((int(o[0]) >> (2 * c)) & 3 for c in range(25)) for o in row) for row in w]


class PredatorPreyModel(ModelFacade):
    def __init__(self, grid: str = "10x10", num_predators: int = 2, num_prey: int = 3, cooperative: bool = True,
                 prey_strength: Optional[int] = None, obs_dim: int = 2, device=None):
        self._init_facade(BatchedPredatorPrey(grid, num_predators, num_prey, cooperative, prey_strength, obs_dim,
                                              device), "PredatorPrey-v0", MAX_EPISODE_STEPS)
        self.action_spaces = {i: Discrete(NUM_ACTIONS) for i in self.possible_agents}
        n_cells = (2 * obs_dim + 1) ** 2
        self.observation_spaces = {
            i: TupleSpace(tuple(Discrete(4) for _ in range(n_cells))) for i in self.possible_agents}
