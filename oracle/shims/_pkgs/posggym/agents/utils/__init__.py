from posggym.agents.utils import processors  # noqa: F401
