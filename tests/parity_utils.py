"""Shared helpers for the parity tests (test infrastructure)."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

STATE_KEYS = ("seg", "ptr", "cash", "long_sh", "short_sh", "margin")


def gbm_ohlc(rng: np.random.Generator, n: int, sigma: float, s0: float = 100.0) -> np.ndarray:
    """Synthetic GBM OHLC bars: r_t~N(0,s^2), C=s0*exp(cumsum r),
    O_t=C_{t-1}*exp(N(0,(s/5)^2)), H/L = max/min(O,C) * exp(+-|N(0,(s/2)^2)|)."""
    r = rng.normal(0.0, sigma, n)
    c = s0 * np.exp(np.cumsum(r))
    prev = np.concatenate([[s0], c[:-1]])
    o = prev * np.exp(rng.normal(0.0, sigma / 5, n))
    o[0] = s0
    h = np.maximum(o, c) * np.exp(np.abs(rng.normal(0.0, sigma / 2, n)))
    l = np.minimum(o, c) * np.exp(-np.abs(rng.normal(0.0, sigma / 2, n)))
    return np.stack([o, h, l, c], axis=1)


def day_labels(num_days: int, bars_per_day, start_minute: int = 9 * 60 + 30):
    """Date/Time strings: each segment is one Date, bar j at 09:30 + j min (<= 390 bars)."""
    import datetime as dt

    if np.isscalar(bars_per_day):
        bars_per_day = [int(bars_per_day)] * num_days
    dates, times = [], []
    d0 = dt.date(2001, 1, 1)
    for d in range(num_days):
        ds = (d0 + dt.timedelta(days=d)).strftime("%m/%d/%Y")
        for j in range(bars_per_day[d]):
            m = start_minute + j
            dates.append(ds)
            times.append(f"{m // 60:02d}:{m % 60:02d}")
    return dates, times


def oracle_state(env) -> dict:
    return {k: getattr(env, k).copy() for k in STATE_KEYS}


def assert_state_equal(a: dict, b: dict, ctx: str = ""):
    for k in STATE_KEYS:
        if not np.array_equal(a[k], b[k], equal_nan=True):
            bad = np.nonzero(~((a[k] == b[k]) | (np.isnan(a[k]) & np.isnan(b[k]))))[0]
            raise AssertionError(f"{ctx}: state '{k}' differs at envs {bad[:8]}: {a[k][bad[:8]]} vs {b[k][bad[:8]]}")


def assert_bits_equal(a: np.ndarray, b: np.ndarray, what: str):
    if not np.array_equal(a, b, equal_nan=True):
        bad = np.argwhere(~((a == b) | (np.isnan(a) & np.isnan(b))))
        raise AssertionError(f"{what}: {len(bad)} mismatches, first at {bad[0]}: {a[tuple(bad[0])]!r} vs {b[tuple(bad[0])]!r}")


# ------------------------------------------------------------------ golden trace replay -----
GOLDEN = os.path.join(ROOT, "tests", "golden")


def golden_traces():
    return sorted(f for f in os.listdir(GOLDEN) if f.endswith(".npz") and f.startswith(("trace_", "kat_")))


def load_trace(name: str):
    with np.load(os.path.join(GOLDEN, name), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}   # materialise once (NpzFile re-inflates on every access)


def trace_params(z) -> dict:
    """(max_shares, starting_balance, commission, imr, mmr) the reference was constructed with (defaults in old traces)."""
    v = z["params"] if "params" in z else np.array([5, 10000.0, 0.01, 1.5, 0.25])
    return {"max_shares": int(v[0]), "starting_balance": float(v[1]), "commission": float(v[2]), "imr": float(v[3]),
            "mmr": float(v[4])}


def trace_series(z):
    from oracle import oracle as orc

    return orc.series_from_prices(z["prices"], z["seg_start"], z["seg_len_raw"], int(z["window"]), logret=z["logret"])


def unpack_state(packed: np.ndarray) -> dict:
    return {
        "seg": packed[0].astype(np.int32), "ptr": packed[1].astype(np.int32), "cash": packed[2].astype(np.float32),
        "long_sh": packed[3].astype(np.float32), "short_sh": packed[4].astype(np.float32), "margin": packed[5],
    }


def replay_trace(z, env, get_state, out_f64: bool, name: str = ""):
    """Drive `env` (oracle or CUDA adapter: reset()/step(a) -> numpy) with a golden trace and compare
    with what the reference produced.  Integer state bit-exact; f32/f64 values bit-exact too (the f32
    outputs of the CUDA path are the reference's f64 values rounded once)."""
    cast = (lambda x: x) if out_f64 else (lambda x: x.astype(np.float32))
    assert_bits_equal(cast(z["obs_reset"]), env.reset(), f"{name} reset obs")
    obs_steps = {int(t): k for k, t in enumerate(z["obs_steps"])}
    info_steps = {int(t): k for k, t in enumerate(z["info_steps"])} if "info_steps" in z else {}
    for t in range(z["actions"].shape[0]):
        o, r, d, info = env.step(z["actions"][t])
        assert_bits_equal(z["dones"][t], d, f"{name} dones t={t}")
        assert_bits_equal(cast(z["rewards"][t]), r, f"{name} rewards t={t}")
        assert_state_equal(unpack_state(z["states"][t]), get_state(), f"{name} t={t}")
        if t in obs_steps:
            assert_bits_equal(cast(z["obs"][obs_steps[t]]), o, f"{name} obs t={t}")
        if out_f64:
            assert_bits_equal(z["obs_sums"][t], o.reshape(o.shape[0], -1).sum(axis=1), f"{name} obs row sums t={t}")
        else:
            np.testing.assert_allclose(o.reshape(o.shape[0], -1).astype(np.float64).sum(axis=1), z["obs_sums"][t],
                                       rtol=1e-5, atol=1e-4, err_msg=f"{name} obs row sums t={t}")
        assert bool(info) == (t in info_steps), f"{name} info presence t={t}"
        if info:
            assert_bits_equal(z["info_returns"][info_steps[t]], np.asarray(info["returns"]), f"{name} returns t={t}")


# ------------------------------------------------------------------ CUDA env adapter --------
class CudaAdapter:
    """Drives finenvs_b200.TimeSeriesEnv with numpy in/out so replay_trace / lock-step loops can
    treat it like the oracle."""

    def __init__(self, env):
        self.env = env

    def reset(self):
        return self.env.reset().cpu().numpy()

    def step(self, actions):
        import torch

        a = torch.from_numpy(np.ascontiguousarray(actions, dtype=np.float32)).view(self.env.num_envs, -1).to(self.env.device)
        o, r, d, info = self.env.step(a)
        return o.cpu().numpy(), r.cpu().numpy(), d.cpu().numpy(), {k: v.cpu().numpy() for k, v in info.items()}

    def state(self):
        e = self.env
        return {
            "seg": e._seg.cpu().numpy(), "ptr": e._ptr.cpu().numpy(), "cash": e._cash.cpu().numpy(),
            "long_sh": e._long.cpu().numpy(), "short_sh": e._short.cpu().numpy(), "margin": e._margin.cpu().numpy(),
        }


def stage_trace_series(z, dtype, device="cuda:0"):
    """Stage a golden trace's series; the log-returns are overwritten with the reference's own values
    (torch.log on CPU) so observations can be compared bit-for-bit; fe_log_returns is tested apart."""
    import torch
    from finenvs_b200.data import loader

    s = loader.stage_series(z["prices"], z["seg_start"], z["seg_len_raw"], int(z["window"]), device, dtype,
                            keep_logret64=True)
    s.logret.copy_(torch.from_numpy(z["logret"]).to(dtype))
    return s
