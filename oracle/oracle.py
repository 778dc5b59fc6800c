"""CPU ORACLE wrapper — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes binding of oracle/fe_oracle.c plus a numpy/pandas restatement of the reference's
loader (hmomin/FinEnvs finenvs/environments/time_series_env.py:80-216).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
finenvs_b200/ never does.

Parity pin: tests/test_oracle_golden.py (golden traces produced by the reference itself,
tests/golden/make_golden.py) and tests/test_oracle_vs_reference.py (live differential run
where /root/reference exists).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libfe_oracle.so")

RESET_KEEP, RESET_LAST, RESET_ALL = 0, 1, 2


class FeoParams(C.Structure):
    _fields_ = [
        ("num_envs", C.c_int64),
        ("env_id_base", C.c_int64),
        ("total_envs", C.c_int64),
        ("num_rows", C.c_int64),
        ("window", C.c_int32),
        ("num_segments", C.c_int32),
        ("num_assets", C.c_int32),
        ("max_shares", C.c_int32),
        ("starting_balance", C.c_double),
        ("commission", C.c_double),
        ("imr", C.c_double),
        ("mmr", C.c_double),
        ("seed", C.c_uint64),
        ("reset_mode", C.c_int32),
        ("random_offset", C.c_int32),
        ("evaluate", C.c_int32),
        ("out_f64", C.c_int32),
    ]


class FeoSeries(C.Structure):
    _fields_ = [
        ("prices", C.c_void_p),
        ("logret", C.c_void_p),
        ("logret32", C.c_void_p),
        ("seg_start", C.c_void_p),
        ("seg_len", C.c_void_p),
    ]


class FeoState(C.Structure):
    _fields_ = [
        ("seg", C.c_void_p),
        ("ptr", C.c_void_p),
        ("cash", C.c_void_p),
        ("long_sh", C.c_void_p),
        ("short_sh", C.c_void_p),
        ("margin", C.c_void_p),
        ("terminated", C.c_void_p),
        ("ep_return", C.c_void_p),
    ]


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc, -ffp-contract=off, OpenMP)."""
    src = os.path.join(_HERE, "fe_oracle.c")
    hdr = os.path.join(_HERE, "fe_oracle.h")
    if (
        not force
        and os.path.exists(_LIB_PATH)
        and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(src), os.path.getmtime(hdr))
    ):
        return _LIB_PATH
    subprocess.check_call(["make", "-s", "-C", _HERE, "-B"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.feo_philox.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.POINTER(C.c_uint32)]
        L.feo_philox.restype = None
        L.feo_draw.argtypes = [C.POINTER(FeoParams), C.POINTER(FeoSeries), C.c_int64, C.c_uint64,
                               C.c_uint32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.feo_draw.restype = None
        L.feo_effective_len.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32]
        L.feo_effective_len.restype = C.c_int32
        L.feo_log_returns.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]
        L.feo_log_returns.restype = None
        L.feo_observe.argtypes = [C.POINTER(FeoParams), C.POINTER(FeoSeries), C.POINTER(FeoState), C.c_void_p]
        L.feo_observe.restype = None
        L.feo_step.argtypes = [C.POINTER(FeoParams), C.POINTER(FeoSeries), C.POINTER(FeoState),
                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                               C.POINTER(C.c_int32)]
        L.feo_step.restype = C.c_int64
        L.feo_observe_multi.argtypes = L.feo_observe.argtypes
        L.feo_observe_multi.restype = None
        L.feo_step_multi.argtypes = L.feo_step.argtypes
        L.feo_step_multi.restype = C.c_int64
        L.feo_reset_all.argtypes = [C.POINTER(FeoParams), C.POINTER(FeoSeries), C.POINTER(FeoState),
                                    C.c_uint64, C.c_int32]
        L.feo_reset_all.restype = None
        L.feo_num_threads.restype = C.c_int
        L.feo_set_num_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def _p(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def philox(seed: int, env_id: int, step: int, kind: int = 0) -> np.ndarray:
    out = (C.c_uint32 * 4)()
    lib().feo_philox(seed, env_id, step, kind, out)
    return np.array(list(out), dtype=np.uint32)


# ----------------------------------------------------------------------------- loader ------
@dataclass
class FlatSeries:
    """Flat series + segment table (the HBM layout, on the host)."""

    prices: np.ndarray      # (T, 4) f64 OHLC, or (T, A, 4) time-major for A assets
    logret: np.ndarray      # same shape, f64
    logret32: np.ndarray    # same shape, f32
    seg_start: np.ndarray   # (D,) i64
    seg_len_raw: np.ndarray  # (D,) i32 = W + bars of the day
    seg_len: np.ndarray     # (D,) i32 effective (NaN probe folded in)
    window: int

    @property
    def num_segments(self) -> int:
        return int(self.seg_start.shape[0])

    @property
    def num_assets(self) -> int:
        return 1 if self.prices.ndim == 2 else int(self.prices.shape[1])

    @property
    def max_len(self) -> int:
        return int(self.seg_len_raw.max())

    def padded(self):
        """Rebuild the reference's NaN-padded (D, L, 4) tensors (:196-216) for loader parity."""
        D, L = self.num_segments, self.max_len
        pe = np.full((D, L, 4), np.nan)
        le = np.full((D, L, 4), np.nan)
        for d in range(D):
            s, n = int(self.seg_start[d]), int(self.seg_len_raw[d])
            pe[d, :n] = self.prices[s:s + n]
            le[d, :n] = self.logret[s:s + n]
        return pe, le


def series_from_prices(prices: np.ndarray, seg_start: np.ndarray, seg_len_raw: np.ndarray, window: int,
                       logret: np.ndarray | None = None) -> FlatSeries:
    prices = np.ascontiguousarray(prices, dtype=np.float64)
    T = prices.shape[0]
    A = 1 if prices.ndim == 2 else prices.shape[1]
    if logret is None:
        logret = np.empty(prices.shape, dtype=np.float64)
        lr32 = np.empty(prices.shape, dtype=np.float32)
        lib().feo_log_returns(_p(prices), T, A, _p(logret), _p(lr32))
    else:
        logret = np.ascontiguousarray(logret, dtype=np.float64)
        lr32 = logret.astype(np.float32)
    seg_start = np.ascontiguousarray(seg_start, dtype=np.int64)
    seg_len_raw = np.ascontiguousarray(seg_len_raw, dtype=np.int32)
    eff = np.array(
        [lib().feo_effective_len(_p(logret), int(s), int(n), window, A) for s, n in zip(seg_start, seg_len_raw)],
        dtype=np.int32,
    )
    return FlatSeries(prices, logret, lr32, seg_start, seg_len_raw, eff, window)


def load_csv(path: str, window: int) -> FlatSeries:
    """Restates read_data :80-88, force_market_hours :90-91, determine_environment_bounds
    :127-152 (day = run of equal Date strings; start = first row - W; days with start < 0 are
    skipped :134), generate_price_dataset :169-177 and generate_log_return_dataset :179-194."""
    import pandas as pd

    df = pd.read_csv(path, names=["Date", "Time", "Open", "High", "Low", "Close", "Volume"])
    df["Datetime"] = pd.to_datetime(df["Date"] + " " + df["Time"])
    df = df.set_index("Datetime").between_time("9:30", "15:59")
    prices = df[["Open", "High", "Low", "Close"]].to_numpy(dtype=np.float64)
    dates = df["Date"].to_numpy()
    seg_start, seg_len = [], []
    # the reference looks up first/last row of each unique Date (:141-152)
    uniq, first = np.unique(dates, return_index=True)
    order = np.argsort(first)
    for u in uniq[order]:
        idx = np.nonzero(dates == u)[0]
        start = int(idx[0]) - window
        if start < 0:
            continue
        seg_start.append(start)
        seg_len.append(int(idx[-1]) - start + 1)
    return series_from_prices(prices, np.array(seg_start), np.array(seg_len), window)


# ------------------------------------------------------------------------------- env -------
class OracleEnv:
    """Host-side env with the same semantics the CUDA path implements.  A = 1 runs the restatement of
    the reference (feo_step); A > 1 (or force_multi) runs the multi-asset extension (feo_step_multi)."""

    def __init__(self, series: FlatSeries, num_envs: int | None = None, *, max_shares=5,
                 starting_balance=10000.0, commission=0.01, imr=1.5, mmr=0.25, evaluate=False,
                 seed=0, reset_mode: int | None = None, random_offset=False, out_f64=True,
                 env_id_base=0, total_envs: int | None = None, seg_init: np.ndarray | None = None,
                 force_multi: bool = False):
        self.series = series
        self.A = A = series.num_assets
        self.multi = force_multi or A > 1
        D = series.num_segments
        if reset_mode is None:
            reset_mode = RESET_KEEP if evaluate else RESET_LAST
        if num_envs is None:
            num_envs = D if evaluate else D + 1      # :246-257
        if total_envs is None:
            total_envs = num_envs
        self.N = N = int(num_envs)
        self.p = FeoParams(N, env_id_base, total_envs, series.prices.shape[0], series.window, D, A,
                           max_shares, starting_balance, commission, imr, mmr, seed, reset_mode,
                           int(random_offset), int(evaluate), int(out_f64))
        self._keep = [series.prices, series.logret, series.logret32, series.seg_start, series.seg_len]
        self.s = FeoSeries(*[_p(a) for a in self._keep])
        self.seg = np.zeros(N, np.int32)
        self.ptr = np.zeros(N, np.int32)
        self.cash = np.full(N, starting_balance, np.float32)
        shp = N if A == 1 else (N, A)
        self.long_sh = np.zeros(shp, np.float32)
        self.short_sh = np.zeros(shp, np.float32)
        self.margin = np.zeros(shp, np.float64)
        self.terminated = np.zeros(N, np.uint8)
        self.ep_return = np.zeros(N, np.float32)
        self.st = FeoState(_p(self.seg), _p(self.ptr), _p(self.cash), _p(self.long_sh), _p(self.short_sh),
                           _p(self.margin), _p(self.terminated), _p(self.ep_return))
        self.step_count = 0
        self.obs_dtype = np.float64 if out_f64 else np.float32
        gid = env_id_base + np.arange(N, dtype=np.int64)
        if seg_init is not None:
            self.seg[:] = seg_init
        else:
            self.seg[:] = gid % D                    # :246 arange(D); widened: i mod D
            if reset_mode == RESET_LAST and env_id_base + N == total_envs:
                # :253-257 the extra (evaluation) env starts on a drawn day
                self.seg[-1] = self.draw(total_envs - 1, 0, 1)[0]
            if reset_mode == RESET_ALL:
                lib().feo_reset_all(C.byref(self.p), C.byref(self.s), C.byref(self.st), 0, 1)

    def draw(self, env_id: int, step: int, kind: int = 0):
        seg, off = C.c_int32(), C.c_int32()
        lib().feo_draw(C.byref(self.p), C.byref(self.s), env_id, step, kind, C.byref(seg), C.byref(off))
        return seg.value, off.value

    def reset(self) -> np.ndarray:
        obs = np.empty((self.N, self.series.window, 5 * self.A), self.obs_dtype)
        fn = lib().feo_observe_multi if self.multi else lib().feo_observe
        fn(C.byref(self.p), C.byref(self.s), C.byref(self.st), _p(obs))
        return obs

    def reset_all(self, redraw: bool | None = None) -> np.ndarray:
        if redraw is None:
            redraw = self.p.reset_mode == RESET_ALL
        lib().feo_reset_all(C.byref(self.p), C.byref(self.s), C.byref(self.st), self.step_count, int(redraw))
        return self.reset()

    def step(self, actions: np.ndarray, want_obs: bool = True):
        actions = np.ascontiguousarray(actions, dtype=np.float32).reshape(self.N * self.A)
        obs = np.empty((self.N, self.series.window, 5 * self.A), self.obs_dtype) if want_obs else None
        rewards = np.empty(self.N, self.obs_dtype)
        dones = np.empty(self.N, np.int32)
        allt = C.c_int32(0)
        self.step_count += 1
        fn = lib().feo_step_multi if self.multi else lib().feo_step
        fn(C.byref(self.p), C.byref(self.s), C.byref(self.st), _p(actions), _p(obs), _p(rewards),
           _p(dones), self.step_count, C.byref(allt))
        info = {}
        if self.p.evaluate and allt.value:           # :531-534
            info = {"returns": self.ep_return.copy()}
            self.ep_return[:] = 0
            self.terminated[:] = 0
        return obs, rewards, dones, info
