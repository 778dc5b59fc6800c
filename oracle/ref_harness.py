"""Differential harness around the REAL reference (TEST INFRASTRUCTURE ONLY).

Imports hmomin/FinEnvs' TimeSeriesEnv straight from the read-only checkout (never copies it),
with a `gym.spaces` shim on sys.path, running on CPU (device_id=-1) from a scratch data
directory so the constructor's `*_bounds_cache.json` side effect (time_series_env.py:154-163)
never touches the checkout.  Its two `torch.randint` call sites (:253 constructor, :511 last-env
redraw) are routed to the oracle's counter-based Philox draw so that reference, oracle and CUDA
kernel consume identical numbers.

Available only where the checkout exists (this container); the GPU box relies on the golden
traces this harness generated (tests/golden/make_golden.py).
"""
from __future__ import annotations

import importlib
import os
import shutil
import sys
import tempfile

import numpy as np

from . import oracle as orc

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("FINENVS_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "finenvs", "environments", "time_series_env.py"))


def data_dir(name: str) -> str:
    return os.path.join(REF_ROOT, "finenvs", "data", name)


_mod = None


def ref_module():
    """The reference module object finenvs.environments.time_series_env."""
    global _mod
    if _mod is None:
        if not available():
            raise RuntimeError(f"reference checkout not found at {REF_ROOT}")
        sys.dont_write_bytecode = True
        shim = os.path.join(_HERE, "_shim")
        for p in (shim, REF_ROOT):
            if p not in sys.path:
                sys.path.insert(0, p)
        _mod = importlib.import_module("finenvs.environments.time_series_env")
    return _mod


class _TorchProxy:
    """Stands in for the `torch` global of the reference module: everything forwards to torch
    except randint, which returns the harness's counter-based draw."""

    def __init__(self, harness):
        import torch

        object.__setattr__(self, "_torch", torch)
        object.__setattr__(self, "_h", harness)

    def __getattr__(self, name):
        if name == "randint":
            return self._h._randint
        return getattr(self._torch, name)


class RefEnv:
    """The reference env + the bookkeeping needed to feed it injected draws."""

    def __init__(self, csv_path: str, key: str = "dummy", window: int = 390, seed: int = 0,
                 evaluate: bool = False, **kwargs):
        import torch

        mod = ref_module()
        self.seed = int(seed)
        self.step_count = 0
        self._in_ctor = True
        self._tmp = tempfile.mkdtemp(prefix="feref_", suffix="_data")  # path must contain "data" (:47-51)
        shutil.copy(csv_path, os.path.join(self._tmp, f"{key}.csv"))
        self._proxy = _TorchProxy(self)
        self.draw_log = []
        mod.torch = self._proxy
        try:
            torch.manual_seed(seed)  # :207 burns global RNG for the NaN padding
            self.env = mod.TimeSeriesEnv(self._tmp, key, num_intervals=window, evaluate=evaluate,
                                         device_id=-1, **kwargs)
        finally:
            self._in_ctor = False
        self.window = window
        self.evaluate = evaluate

    def close(self):
        ref_module().torch = self._proxy._torch
        shutil.rmtree(self._tmp, ignore_errors=True)

    # -- injected RNG -------------------------------------------------------------------
    def _randint(self, low, high, size, device=None, **kw):
        import torch

        # constructor (:253): the extra env gets id D == high (arange(D) precede it, :246-257);
        # step time (:511): the redrawn env is always the last one.
        env_id = int(high) if self._in_ctor else int(self.env.num_envs - 1)
        kind = 1 if self._in_ctor else 0
        r = orc.philox(self.seed, env_id, self.step_count, kind)
        seg = (int(r[0]) * int(high - low)) >> 32
        self.draw_log.append((self.step_count, kind, seg))
        return torch.tensor([low + seg], dtype=torch.int64)

    # -- widening -------------------------------------------------------
    def widen(self, seg_init: np.ndarray):
        import torch

        e = self.env
        N = int(len(seg_init))
        e.env_indices = torch.as_tensor(np.asarray(seg_init), dtype=torch.int64)
        e.num_envs = N
        e.env_pointers = torch.zeros((N,), dtype=torch.int64)
        e.env_spots = torch.arange(0, e.num_intervals).repeat(N, 1)
        e.cash = e.starting_balance * torch.ones((N, 1))
        e.long_shares = torch.zeros((N, 1))
        e.short_shares = torch.zeros((N, 1))
        e.margin = torch.zeros((N, 1))
        if self.evaluate:
            e.reset_evaluation_metrics()
        if hasattr(e, "current_close_prices"):
            del e.current_close_prices

    # -- stepping ---------------------------------------------------------------------------
    def reset(self) -> np.ndarray:
        return self.env.reset().numpy().copy()

    def step(self, actions: np.ndarray):
        import torch

        self.step_count += 1
        a = torch.as_tensor(np.asarray(actions, dtype=np.float32)).view(-1, 1).clone()
        obs, rew, dones, info = self.env.step(a)
        info = {k: v.numpy().copy() for k, v in info.items()}
        return obs.numpy().copy(), rew.numpy().copy(), dones.numpy().copy(), info

    def state(self) -> dict:
        e = self.env
        d = {
            "seg": e.env_indices.numpy().astype(np.int32).copy(),
            "ptr": e.env_pointers.numpy().astype(np.int32).copy(),
            "cash": e.cash.numpy().reshape(-1).copy(),
            "long_sh": e.long_shares.numpy().reshape(-1).copy(),
            "short_sh": e.short_shares.numpy().reshape(-1).copy(),
            "margin": e.margin.numpy().reshape(-1).astype(np.float64).copy(),
        }
        assert d["cash"].dtype == np.float32
        # :261/:282/:521 env_spots[i, j] must stay ptr[i] + j — the flat layout relies on it
        spots = e.env_spots.numpy()
        assert (spots == d["ptr"][:, None] + np.arange(spots.shape[1])[None, :]).all()
        return d


def flat_series_from_ref(ref: RefEnv) -> orc.FlatSeries:
    """Flat layout rebuilt from the reference's own tensors (dataset + cached bounds)."""
    e = ref.env
    import json

    with open(os.path.join(ref._tmp, f"{e.file_key}_bounds_cache.json")) as f:
        cache = json.load(f)
    starts = np.array(cache["start_indices"], dtype=np.int64)
    stops = np.array(cache["stop_indices"], dtype=np.int64)
    return orc.series_from_prices(e.dataset.numpy(), starts, (stops - starts + 1).astype(np.int32),
                                  ref.window, logret=e.log_return_dataset.numpy())


def write_csv(path: str, dates, times, ohlc: np.ndarray, volume=None, decimals: int = 4):
    """Reference-format headerless CSV (finenvs/data/README.md:7-11)."""
    with open(path, "w") as f:
        for i in range(len(dates)):
            o, h, l, c = ohlc[i]
            v = 0 if volume is None else int(volume[i])
            f.write(f"{dates[i]},{times[i]},{o:.{decimals}f},{h:.{decimals}f},{l:.{decimals}f},{c:.{decimals}f},{v}\n")
