"""The UNMODIFIED reference (hmomin/FinEnvs) as an importable tree under baseline/_ref/ — baseline and
compatibility-test infrastructure only; nothing under finenvs_b200/ imports this.

The reference ships no setup.py / pyproject.toml, so `pip install --target baseline/_ref /root/reference`
cannot work ("neither 'setup.py' nor 'pyproject.toml' found"); `install()` copies the package directory as it
is (minus the 100 MB Isaac Gym asset tree, which needs the proprietary isaacgym binary and is out of scope)
next to the 8-line `gym.spaces` stand-in the reference's import needs (gym is not in this image).
baseline/_ref/ is git-ignored (no reference source enters the history) but travels to the GPU box.

Used by: bench.py (`--impl reference` and the `cpu_baseline` leg: the reference's own TimeSeriesEnv.step timed
on the host cores), tests/test_gpu_reference_agents.py (the reference's PPO / ES agents stepping the drop-in env).
"""
from __future__ import annotations

import importlib
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = os.environ.get("FINENVS_REFERENCE", "/root/reference")
REF_DST = os.path.join(ROOT, "baseline", "_ref")
_SHIM = os.path.join(ROOT, "oracle", "_shim", "gym")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DST, "finenvs", "environments", "time_series_env.py"))


def install(force: bool = False) -> bool:
    """Copy the reference package into baseline/_ref (only where the checkout exists). True = installed."""
    src = os.path.join(REF_SRC, "finenvs")
    if not os.path.isdir(src):
        return available()
    if available() and not force:
        return True
    if os.path.isdir(REF_DST):
        shutil.rmtree(REF_DST)
    os.makedirs(REF_DST)
    shutil.copytree(src, os.path.join(REF_DST, "finenvs"),
                    ignore=shutil.ignore_patterns("isaac_gym_envs", "__pycache__", "*_bounds_cache.json"))
    shutil.copytree(_SHIM, os.path.join(REF_DST, "gym"), ignore=shutil.ignore_patterns("__pycache__"))
    return True


def import_reference():
    """Put baseline/_ref on sys.path and return the reference's `finenvs` package."""
    if not available():
        raise RuntimeError("baseline/_ref is missing: run `python __graft_entry__.py` where /root/reference exists")
    if REF_DST not in sys.path:
        sys.path.insert(0, REF_DST)
    return importlib.import_module("finenvs")


def data_dir(name: str) -> str:
    return os.path.join(REF_DST, "finenvs", "data", name)
