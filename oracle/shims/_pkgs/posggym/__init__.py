"""Minimal stand-in for posggym (oracle shim).  ``make`` builds the oracle's
restated models wrapped in a tiny Env (reset/step/TimeLimit of 50)."""
from posggym import model  # noqa: F401
from posggym.model import POSGModel  # noqa: F401
from posggym.core import Env, make  # noqa: F401
