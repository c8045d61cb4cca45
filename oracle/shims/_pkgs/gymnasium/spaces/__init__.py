"""gymnasium.spaces stand-in: Discrete / Tuple / Box with the attributes the
reference reads (``.n``, ``.sample()``, ``.shape``; search_policy.py:163-180,
mcts.py:46-48, intmcp.py:58-64).  ``Discrete.sample`` draws from a numpy
PCG64 generator like gymnasium 0.29 does."""
from __future__ import annotations

import numpy as np


class Space:
    def __init__(self, shape=None, dtype=None, seed=None):
        self._shape = shape
        self.dtype = dtype
        self._np_random = None
        self._seed = seed

    @property
    def shape(self):
        return self._shape

    @property
    def np_random(self):
        if self._np_random is None:
            self._np_random = np.random.Generator(np.random.PCG64(self._seed))
        return self._np_random

    def seed(self, seed=None):
        self._seed = seed
        self._np_random = np.random.Generator(np.random.PCG64(seed))
        return [seed]


class Discrete(Space):
    def __init__(self, n, seed=None, start=0):
        super().__init__((), np.int64, seed)
        self.n = int(n)
        self.start = int(start)

    def sample(self, mask=None):
        return int(self.start + self.np_random.integers(self.n))

    def contains(self, x):
        return isinstance(x, (int, np.integer)) and self.start <= int(x) < self.start + self.n

    __contains__ = contains

    def __repr__(self):
        return f"Discrete({self.n})"

    def __eq__(self, other):
        return isinstance(other, Discrete) and self.n == other.n and self.start == other.start


class Tuple(Space):
    def __init__(self, spaces, seed=None):
        super().__init__(None, None, seed)
        self.spaces = tuple(spaces)

    def sample(self, mask=None):
        return tuple(s.sample() for s in self.spaces)

    def __len__(self):
        return len(self.spaces)

    def __getitem__(self, i):
        return self.spaces[i]

    def __iter__(self):
        return iter(self.spaces)

    def __repr__(self):
        return "Tuple(" + ", ".join(repr(s) for s in self.spaces) + ")"


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
        if shape is None:
            shape = np.asarray(low).shape
        super().__init__(tuple(shape), dtype, seed)
        self.low = np.full(shape, low, dtype=dtype) if np.isscalar(low) else np.asarray(low, dtype)
        self.high = np.full(shape, high, dtype=dtype) if np.isscalar(high) else np.asarray(high, dtype)

    def sample(self, mask=None):
        return self.np_random.uniform(self.low, self.high).astype(self.dtype)


def flatdim(space) -> int:
    if isinstance(space, Discrete):
        return space.n
    if isinstance(space, Tuple):
        return sum(flatdim(s) for s in space.spaces)
    if isinstance(space, Box):
        return int(np.prod(space.shape))
    raise NotImplementedError(type(space))


def flatten(space, x) -> np.ndarray:
    """gymnasium.spaces.flatten: one-hot for Discrete, concatenation for Tuple."""
    if isinstance(space, Discrete):
        out = np.zeros(space.n, dtype=np.int64)
        out[int(x) - space.start] = 1
        return out
    if isinstance(space, Tuple):
        return np.concatenate([flatten(s, xi) for s, xi in zip(space.spaces, x)])
    if isinstance(space, Box):
        return np.asarray(x, dtype=space.dtype).reshape(-1)
    raise NotImplementedError(type(space))


def flatten_space(space) -> Box:
    if isinstance(space, Box):
        return Box(space.low.reshape(-1), space.high.reshape(-1), dtype=space.dtype)
    return Box(0, 1, (flatdim(space),), dtype=np.int64)
