from .buffer import Buffer  # noqa: F401
