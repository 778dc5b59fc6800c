from .time_series_env import TimeSeriesEnv  # noqa: F401
