"""Minimal stand-in for the `gym` package (TEST INFRASTRUCTURE ONLY).

The reference env imports `gym.spaces` purely for `Box` metadata
(finenvs/environments/time_series_env.py:9, :224-234); gym is not installed in
this image, so the differential harness puts this shim on sys.path before
importing the reference.
"""
from . import spaces  # noqa: F401
