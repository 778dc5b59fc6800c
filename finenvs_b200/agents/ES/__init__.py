from .evo_agent import EvoAgent  # noqa: F401
