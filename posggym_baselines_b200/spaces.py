"""Action/observation space objects of the model facade.

Uses the real ``gymnasium.spaces`` when it is importable (so that the reference's
``isinstance(space, gym.spaces.Discrete)`` check, planning/mcts.py:46, passes) and a
small stand-in with the same attributes (``n``, ``sample()``, ``seed()``) otherwise.
"""
from __future__ import annotations

try:  # pragma: no cover - depends on the environment
    from gymnasium.spaces import Discrete, Tuple  # type: ignore  # noqa: F401
except Exception:  # gymnasium absent
    import numpy as _np

    class Discrete:  # type: ignore[no-redef]
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

    class Tuple:  # type: ignore[no-redef]
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
