"""finenvs_b200 — the vectorised trading-env step of hmomin/FinEnvs as one sm_100a CUDA kernel.

Drop-in surface: `finenvs_b200.environments.TimeSeriesEnv` keeps the reference's
`reset()` / `step(actions)` / `get_env_args()` interface (torch tensors in and out).
"""
from . import _lib  # noqa: F401  (binding only; the .so is loaded on first use and its absence is fatal)
from .device_utils import set_device  # noqa: F401

__all__ = ["set_device", "TimeSeriesEnv"]


def __getattr__(name):
    if name == "TimeSeriesEnv":
        from .environments.time_series_env import TimeSeriesEnv

        return TimeSeriesEnv
    raise AttributeError(name)
