from .parallel_mlp import ParallelMLP  # noqa: F401
