"""Minimal stand-in for gymnasium (oracle shim; only what the reference touches)."""
from gymnasium import spaces  # noqa: F401
