from .time_series_env import CapturedRollout, LazyObs, TimeSeriesEnv  # noqa: F401
