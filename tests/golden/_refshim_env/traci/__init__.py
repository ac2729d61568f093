"""Import-only stand-in for SUMO's `traci` (absent): nothing here is ever called."""
from . import exceptions  # noqa: F401


def close(*args, **kwargs):
    pass
