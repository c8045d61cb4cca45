"""Action/observation space objects of the model facade.

``Discrete(...)`` / ``Tuple(...)`` return the real ``gymnasium.spaces`` objects whenever a ``gymnasium``
module is importable AT CALL TIME (so that the reference's ``isinstance(space, gym.spaces.Discrete)``
check, planning/mcts.py:46, passes for a model built after gymnasium -- or a stand-in for it -- became
importable) and a small stand-in with the same attributes (``n``, ``sample()``, ``seed()``) otherwise.
"""
from __future__ import annotations

import sys

import numpy as _np


def _gym_spaces():
    mod = sys.modules.get("gymnasium")
    if mod is None:
        try:
            import gymnasium as mod  # type: ignore  # noqa: F811
        except Exception:
            return None
    return getattr(mod, "spaces", None)


class _Discrete:
    def __init__(self, n: int, seed=None, start: int = 0):
        self.n = int(n)
        self.start = int(start)
        self._rng = _np.random.Generator(_np.random.PCG64(seed))

    def seed(self, seed=None):
        self._rng = _np.random.Generator(_np.random.PCG64(seed))
        return [seed]

    def sample(self, mask=None) -> int:
        return int(self.start + self._rng.integers(self.n))

    def contains(self, x) -> bool:
        return self.start <= int(x) < self.start + self.n

    __contains__ = contains

    def __repr__(self):
        return f"Discrete({self.n})"


class _Tuple:
    def __init__(self, spaces, seed=None):
        self.spaces = tuple(spaces)

    def sample(self, mask=None):
        return tuple(s.sample() for s in self.spaces)

    def __len__(self):
        return len(self.spaces)

    def __getitem__(self, i):
        return self.spaces[i]

    def __repr__(self):
        return "Tuple(" + ", ".join(map(repr, self.spaces)) + ")"


def Discrete(n: int, seed=None, start: int = 0):  # noqa: N802 - mirrors the class name
    sp = _gym_spaces()
    if sp is not None:
        return sp.Discrete(n, seed=seed, start=start) if start else sp.Discrete(n, seed=seed)
    return _Discrete(n, seed=seed, start=start)


def Tuple(spaces, seed=None):  # noqa: N802
    sp = _gym_spaces()
    if sp is not None:
        return sp.Tuple(tuple(spaces), seed=seed) if seed is not None else sp.Tuple(tuple(spaces))
    return _Tuple(spaces, seed=seed)
