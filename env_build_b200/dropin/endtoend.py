"""Flat-module shim for `from endtoend import CrossroadEnd2end` (reference hier_decision.py:22,
multi_ego.py:21): the batched, SUMO-free environment of env_build_b200.endtoend."""
from env_build_b200.endtoend import CrossroadEnd2end  # noqa: F401
