"""Import shims (oracle; test infrastructure only).

``install()`` makes ``import gymnasium`` / ``import posggym`` resolve to the
minimal stand-ins under ``oracle/shims/_pkgs`` when the real packages are
absent (they are absent in the build container and on the GPU box: SURVEY.md
section 0 fact 3), and fabricates empty stub modules for the plotting / logging
packages the reference imports at module top but never touches on the planning
path (matplotlib, seaborn, wandb, pygame).  With the shims installed the
UNMODIFIED reference planner (``/root/reference/posggym_baselines/planning``)
imports and drives any duck-typed ``POSGModel`` -- the API surface needed is the
one enumerated in SURVEY.md section 8(b).1 and 8(b).5.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys
import types

_PKGS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_pkgs")
_STUBBED = ("matplotlib", "seaborn", "wandb", "pygame", "tensorboard")


class _Anything:
    """Permissive placeholder: callable, attribute-able, usable as a base class."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def __init__(self, roots):
        self.roots = tuple(roots)

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in self.roots:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


def _importable(name: str) -> bool:
    try:
        return importlib.util.find_spec(name) is not None
    except (ImportError, ValueError):
        return False


_installed = False


def install(reference_root: str = "/root/reference") -> bool:
    """Install the shims.  Returns True when the reference tree is importable."""
    global _installed
    if not _installed:
        if not (_importable("posggym") and _importable("gymnasium")):
            if _PKGS not in sys.path:
                sys.path.insert(0, _PKGS)
        missing = [r for r in _STUBBED if not _importable(r)]
        if "torch" in sys.modules or _importable("torch"):
            # torch.utils.tensorboard needs the real tensorboard; stub the torch side too
            pass
        if missing:
            sys.meta_path.append(_StubFinder(missing))
        _installed = True
    have_ref = os.path.isdir(os.path.join(reference_root, "posggym_baselines"))
    if have_ref and reference_root not in sys.path:
        sys.path.append(reference_root)
    return have_ref
