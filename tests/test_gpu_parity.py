"""GPU parity tests proper: the CUDA path (through the C ABI) against the golden traces the
reference produced and against the CPU oracle on seeded inputs.  Bit-exact everywhere: integers by
construction, floats because the kernel performs the reference's IEEE operations in the reference's
order (the stated tolerance of 1e-5 relative is therefore met with margin 0).

Where a test builds its series with `keep_logret64=True` and hands `series.logret64` to the oracle, the oracle steps on
the log-return table THE GPU COMPUTED (fe_log_returns; <= 2 ulp from torch.log, pinned separately by
test_log_returns_kernel_vs_reference_values): those tests pin the step, not the table.  The golden-trace tests inject
the reference's own table instead."""
import numpy as np
import pytest
import torch

from parity_utils import (CudaAdapter, assert_bits_equal, assert_state_equal, gbm_ohlc, golden_traces, load_trace, trace_params,
                          oracle_state, replay_trace, stage_trace_series)

pytestmark = pytest.mark.gpu

RTOL = 1e-5  # north-star tolerance for fp32 cash / position / reward / observation values


def _lib_mod():
    from finenvs_b200 import _lib

    return _lib


def _env(series, **kw):
    from finenvs_b200.environments import TimeSeriesEnv

    return TimeSeriesEnv("golden", num_intervals=series.window, device_id=0, series=series, **kw)


@pytest.mark.parametrize("name", golden_traces())
@pytest.mark.parametrize("dtype,variant", [(torch.float32, "auto"), (torch.float64, "auto"), (torch.float32, "direct"),
                                           (torch.float64, "direct"), (torch.float32, "portfolio"),
                                           (torch.float64, "portfolio"), (torch.float32, "pipe"),
                                           (torch.float64, "pipe"), (torch.float32, "gather"),
                                           (torch.float64, "gather"), (torch.float32, "split"),
                                           (torch.float64, "split"), (torch.float32, "tile"), (torch.float64, "tile")])
def test_cuda_replays_reference_trace(name, dtype, variant):
    z = load_trace(name)
    if variant == "pipe" and _lib_mod().lib().fe_pipe_envs(int(z["window"]), int(dtype == torch.float64), 0) == 0:
        pytest.skip("window too large for the pipe variant's rings")
    if variant == "gather" and _lib_mod().lib().fe_obs_table_bytes(int(z["prices"].shape[0]), int(z["window"]), int(dtype == torch.float64)) == 0:
        pytest.skip("this window has no gather variant (5*W*itemsize a multiple of 16 and <= 2048 bytes, or two such halves)")
    if variant == "tile" and _lib_mod().lib().fe_tile_envs(int(z["window"]), int(dtype == torch.float64), 0) == 0:
        pytest.skip("window too large for the tile variant")
    series = stage_trace_series(z, dtype)
    N = len(z["seg_init"])
    tp = trace_params(z)
    env = _env(series, num_envs=N, evaluate=bool(z["evaluate"]), seed=int(z["seed"]), obs_dtype=dtype, variant=variant,
               max_shares=tp["max_shares"], starting_balance=tp["starting_balance"], per_share_commission=tp["commission"],
               initial_margin_requirement=tp["imr"], maintenance_margin_requirement=tp["mmr"])
    env._seg.copy_(torch.from_numpy(z["seg_init"]))
    ad = CudaAdapter(env)
    replay_trace(z, ad, ad.state, dtype == torch.float64, f"{name}[{variant}]")


def _c1_series(W=60, days=1024, bars=252, sigma=0.01, seed=20260101):
    from finenvs_b200.data import loader

    rng = np.random.default_rng(seed)
    prices = np.round(gbm_ohlc(rng, days * bars, sigma), 4)
    seg_start, seg_len = loader.regular_segments(days * bars, bars, W)
    return prices, seg_start, seg_len


def _lockstep(env, ref, steps, rng, obs_every=1, action_fn=None):
    ad = CudaAdapter(env)
    assert_state_equal(oracle_state(ref), ad.state(), "init")
    assert_bits_equal(ref.reset(), ad.reset(), "reset obs")
    n_done = 0
    for t in range(steps):
        a = action_fn(rng, ref.N) if action_fn else rng.uniform(-1, 1, ref.N).astype(np.float32)
        want = (t % obs_every) == 0
        o_ref, r_ref, d_ref, i_ref = ref.step(a, want_obs=want)
        o, r, d, info = ad.step(a)
        assert_bits_equal(d_ref, d, f"dones t={t}")
        assert_state_equal(oracle_state(ref), ad.state(), f"t={t}")
        np.testing.assert_allclose(r, r_ref, rtol=RTOL, atol=0, err_msg=f"rewards t={t}")
        assert_bits_equal(r_ref, r, f"rewards t={t}")
        if want:
            np.testing.assert_allclose(o, o_ref, rtol=RTOL, atol=0, err_msg=f"obs t={t}")
            assert_bits_equal(o_ref, o, f"obs t={t}")
        assert set(info) == set(i_ref)
        n_done += int(d_ref.sum())
    return n_done


@pytest.mark.parametrize("mode,offset", [("last", False), ("all", True), ("keep", False)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_config1_1024_envs_vs_oracle(mode, offset, dtype):
    """BASELINE config 1: single asset, GBM daily bars, 1024 envs, W=60, random actions; 2x252+ steps
    so every env auto-resets at least twice."""
    from oracle import oracle as orc
    from finenvs_b200.data import loader

    W = 60
    prices, seg_start, seg_len = _c1_series(W)
    series = loader.stage_series(prices, seg_start, seg_len, W, "cuda:0", dtype, keep_logret64=True)
    assert series.num_segments == 1023
    env = _env(series, seed=1234, random_reset=mode, random_offset=offset, obs_dtype=dtype)  # N = D + 1 = 1024
    assert env.num_envs == 1024
    fs = orc.series_from_prices(prices, seg_start, seg_len, W, logret=series.logret64.cpu().numpy())
    ref = orc.OracleEnv(fs, num_envs=1024, seed=1234, reset_mode={"last": 1, "all": 2, "keep": 0}[mode],
                        random_offset=offset, out_f64=dtype == torch.float64)
    n_done = _lockstep(env, ref, 520, np.random.default_rng(1234), obs_every=7)
    assert n_done >= 2 * 1024


def test_adversarial_branches_vs_oracle():
    """sigma=0.15 bars and short-biased actions: margin calls at High and Close, releases, bankruptcies,
    blocked entries; ragged segment lengths; windows W=8 (tile) at N not a multiple of the tile."""
    from oracle import oracle as orc
    from finenvs_b200.data import loader

    rng = np.random.default_rng(5)
    bars = rng.integers(1, 60, 200)
    W = 8
    prices = np.round(gbm_ohlc(rng, int(bars.sum()) + W, 0.15), 4)
    firsts = W + np.concatenate([[0], np.cumsum(bars)[:-1]])
    seg_start, seg_len = firsts - W, (bars + W).astype(np.int32)
    series = loader.stage_series(prices, seg_start, seg_len, W, "cuda:0", torch.float32, keep_logret64=True)
    N = 5003
    env = _env(series, num_envs=N, seed=9, random_reset="all", random_offset=True)
    fs = orc.series_from_prices(prices, seg_start, seg_len, W, logret=series.logret64.cpu().numpy())
    ref = orc.OracleEnv(fs, num_envs=N, seed=9, reset_mode=2, random_offset=True, out_f64=False)

    def act(r, n):
        a = r.uniform(-1, 1, n)
        m = r.uniform(0, 1, n) < 0.6
        a[m] = -np.abs(a[m])
        return a.astype(np.float32)

    n_done = _lockstep(env, ref, 300, rng, obs_every=3, action_fn=act)
    assert n_done > 10000


@pytest.mark.parametrize("W", [1, 3, 4, 61, 390, 1500, 2000])
def test_window_sizes_and_variants(W):
    """Odd windows (tail blocks whose byte count is not a multiple of 16), the reference default 390,
    the largest windows the tile variant takes and one that must fall back to the direct variant."""
    from oracle import oracle as orc
    from finenvs_b200 import _lib
    from finenvs_b200.data import loader

    rng = np.random.default_rng(W)
    bars, days = 20, 12
    T = W + bars * days
    prices = np.round(gbm_ohlc(rng, T, 0.02), 4)
    firsts = W + bars * np.arange(days)
    seg_start, seg_len = firsts - W, np.full(days, W + bars, np.int32)
    for dtype in (torch.float32, torch.float64):
        series = loader.stage_series(prices, seg_start, seg_len, W, "cuda:0", dtype, keep_logret64=True)
        fs = orc.series_from_prices(prices, seg_start, seg_len, W, logret=series.logret64.cpu().numpy())
        tile = _lib.lib().fe_tile_envs(W, int(dtype == torch.float64), 0)
        for variant in (["auto", "direct"] if tile else ["auto"]):
            N = 37
            env = _env(series, num_envs=N, seed=3, random_reset="all", obs_dtype=dtype, variant=variant)
            ref = orc.OracleEnv(fs, num_envs=N, seed=3, reset_mode=2, out_f64=dtype == torch.float64)
            _lockstep(env, ref, 45, np.random.default_rng(W + 1))
    if W == 2000:
        assert _lib.lib().fe_tile_envs(W, 0, 0) == 0
        series32 = loader.stage_series(prices, seg_start, seg_len, W, "cuda:0", torch.float32)
        with pytest.raises(_lib.FeError):
            _env(series32, num_envs=8, variant="tile").reset()


def test_log_returns_kernel_vs_reference_values():
    """fe_log_returns (:179-194) against the reference's own table (torch.log on CPU): <= 2 ulp."""
    from finenvs_b200.data import loader

    for name in ("trace_oih_w60.npz", "trace_spy_w390.npz", "trace_adv_s15_w8_n96.npz"):
        z = load_trace(name)
        s = loader.stage_series(z["prices"], z["seg_start"], z["seg_len_raw"], int(z["window"]), "cuda:0", torch.float32,
                                keep_logret64=True)
        got = s.logret64.cpu().numpy()
        np.testing.assert_allclose(got, z["logret"], rtol=4e-16, atol=2e-14)
        np.testing.assert_allclose(s.logret.cpu().numpy(), z["logret"].astype(np.float32), rtol=2e-7, atol=1e-12)
        assert np.array_equal(s.seg_len.cpu().numpy(), s.seg_len_raw.cpu().numpy())


def test_nan_rows_end_the_segment_like_the_reference_probe():
    """A zero price makes log() NaN; the reference then treats the row as end-of-day padding (:486-496)."""
    from oracle import oracle as orc
    from finenvs_b200.data import loader

    rng = np.random.default_rng(2)
    W, bars, days = 4, 30, 6
    prices = np.round(gbm_ohlc(rng, W + bars * days, 0.01), 4)
    prices[W + 2 * bars + 10, 0] = -1.0   # log(neg/..) -> NaN in col 0 of that row ... and col 1-3 NaN too
    firsts = W + bars * np.arange(days)
    seg_start, seg_len = firsts - W, np.full(days, W + bars, np.int32)
    series = loader.stage_series(prices, seg_start, seg_len, W, "cuda:0", torch.float64, keep_logret64=True)
    fs = orc.series_from_prices(prices, seg_start, seg_len, W, logret=series.logret64.cpu().numpy())
    assert np.array_equal(series.seg_len.cpu().numpy(), fs.seg_len)
    assert fs.seg_len[2] == W + 10 and (fs.seg_len[[0, 1, 3, 4, 5]] == W + bars).all()
    env = _env(series, num_envs=days, seed=1, random_reset="keep", obs_dtype=torch.float64)
    ref = orc.OracleEnv(fs, num_envs=days, seed=1, reset_mode=0, out_f64=True)
    _lockstep(env, ref, 70, rng)


def test_evaluate_mode_returns_and_stats():
    from oracle import oracle as orc
    from finenvs_b200.data import loader

    W = 16
    prices, seg_start, seg_len = _c1_series(W, days=64, bars=40, sigma=0.03)
    series = loader.stage_series(prices, seg_start, seg_len, W, "cuda:0", torch.float32, keep_logret64=True)
    fs = orc.series_from_prices(prices, seg_start, seg_len, W, logret=series.logret64.cpu().numpy())
    env = _env(series, evaluate=True, seed=2)
    assert env.num_envs == series.num_segments
    ref = orc.OracleEnv(fs, evaluate=True, seed=2, out_f64=False)
    ad = CudaAdapter(env)
    rng = np.random.default_rng(0)
    n_info = 0
    for t in range(130):
        a = rng.uniform(-1, 1, ref.N).astype(np.float32)
        o_ref, r_ref, d_ref, i_ref = ref.step(a)
        o, r, d, info = ad.step(a)
        assert_bits_equal(r_ref, r, f"rewards t={t}")
        assert_bits_equal(d_ref, d, f"dones t={t}")
        assert set(info) == set(i_ref)
        if info:
            n_info += 1
            assert_bits_equal(i_ref["returns"], info["returns"], "returns")
        assert_bits_equal(ref.ep_return, env.episode_returns.cpu().numpy(), f"ep_return t={t}")
    assert n_info == 3


def test_track_stats_matches_host_accounting():
    from finenvs_b200.data import loader

    W = 16
    prices, seg_start, seg_len = _c1_series(W, days=64, bars=40, sigma=0.03)
    series = loader.stage_series(prices, seg_start, seg_len, W, "cuda:0", torch.float32)
    N = 3000
    env = _env(series, num_envs=N, seed=2, random_reset="all", random_offset=True, track_stats=True)
    g = torch.Generator().manual_seed(0)
    ep_ret = np.zeros(N)
    ep_len = np.zeros(N, np.int64)
    fin_ret, fin_len = [], []
    for t in range(100):
        a = (torch.rand((N, 1), generator=g) * 2 - 1).cuda()
        _, r, d, _ = env.step(a)
        r, d = r.cpu().numpy().astype(np.float64), d.cpu().numpy().astype(bool)
        ep_ret = (ep_ret + r).astype(np.float32).astype(np.float64)
        ep_len += 1
        fin_ret += list(ep_ret[d]); fin_len += list(ep_len[d])
        ep_ret[d] = 0; ep_len[d] = 0
    st = {k: v.item() for k, v in env.stats().items()}
    assert st["n_done"] == len(fin_ret) > 0
    assert st["sum_len"] == int(np.sum(fin_len))
    np.testing.assert_allclose(st["sum_return"], np.sum(fin_ret), rtol=1e-6)  # host side re-rounds the f32 rewards
    np.testing.assert_allclose(st["sum_return_sq"], np.sum(np.square(fin_ret)), rtol=1e-6)


def test_sharding_does_not_change_results():
    """Env-sharded run (two shards of one population, as two GPUs would hold) == unsharded run."""
    from finenvs_b200.data import loader

    W = 16
    prices, seg_start, seg_len = _c1_series(W, days=64, bars=40, sigma=0.03)
    series = loader.stage_series(prices, seg_start, seg_len, W, "cuda:0", torch.float32)
    N, cut = 2048, 1100
    kw = dict(seed=77, random_reset="all", random_offset=True)
    full = _env(series, num_envs=N, **kw)
    lo = _env(series, num_envs=cut, env_id_base=0, total_envs=N, **kw)
    hi = _env(series, num_envs=N - cut, env_id_base=cut, total_envs=N, **kw)
    g = torch.Generator().manual_seed(5)
    for t in range(90):
        a = (torch.rand((N, 1), generator=g) * 2 - 1).cuda()
        o, r, d, _ = full.step(a)
        o1, r1, d1, _ = lo.step(a[:cut].contiguous())
        o2, r2, d2, _ = hi.step(a[cut:].contiguous())
        assert torch.equal(o, torch.cat([o1, o2])) and torch.equal(r, torch.cat([r1, r2])) and torch.equal(d, torch.cat([d1, d2]))
        assert torch.equal(full._seg, torch.cat([lo._seg, hi._seg])) and torch.equal(full._ptr, torch.cat([lo._ptr, hi._ptr]))
    # reference-style "last env" redraw lives on the shard that owns the global last id
    last_full = _env(series, num_envs=N, seed=3, random_reset="last")
    last_hi = _env(series, num_envs=N - cut, env_id_base=cut, total_envs=N, seed=3, random_reset="last")
    assert int(last_full._seg[-1]) == int(last_hi._seg[-1])


# ------------------------------------------------------------------ portfolio (A > 1) extension ----
def _portfolio_series(A, W, days, bars, sigma, seed):
    from finenvs_b200.data import loader

    rng = np.random.default_rng(seed)
    T = days * bars
    prices = np.stack([np.round(gbm_ohlc(rng, T, sigma, s0=20.0 + 7 * a), 4) for a in range(A)], axis=1)  # (T, A, 4)
    seg_start, seg_len = loader.regular_segments(T, bars, W)
    return prices, seg_start, seg_len


@pytest.mark.parametrize("A,W,N,sigma,dtype", [(30, 128, 384, 0.01, torch.float32), (30, 128, 96, 0.01, torch.float64),
                                               (5, 7, 1001, 0.1, torch.float32), (3, 5, 257, 0.12, torch.float64),
                                               (32, 16, 200, 0.05, torch.float32), (2, 60, 333, 0.08, torch.float32)])
def test_portfolio_env_vs_oracle(A, W, N, sigma, dtype):
    """BASELINE config 3 shape (30 assets, W=128, transaction costs) at a lock-step-able env count, plus small /
    odd shapes (unaligned chunk tails -> plain-store path, A = 32 full warp, high sigma -> margin calls and
    bankruptcies across assets).  No reference exists for A > 1: the oracle's multi-asset code is the one
    proven equal to the reference at A = 1 (tests/test_oracle_golden.py)."""
    from oracle import oracle as orc
    from finenvs_b200.data import loader

    bars = 40
    prices, seg_start, seg_len = _portfolio_series(A, W, days=10 + W // bars, bars=bars, sigma=sigma, seed=A * 1000 + W)
    series = loader.stage_series(prices, seg_start, seg_len, W, "cuda:0", dtype, keep_logret64=True)
    assert series.num_assets == A
    env = _env(series, num_envs=N, seed=5, random_reset="all", random_offset=True, obs_dtype=dtype, track_stats=True)
    assert env.num_acts == A and env.num_obs == 5 * A
    fs = orc.series_from_prices(prices, seg_start, seg_len, W, logret=series.logret64.cpu().numpy())
    ref = orc.OracleEnv(fs, num_envs=N, seed=5, reset_mode=2, random_offset=True, out_f64=dtype == torch.float64)
    rng = np.random.default_rng(A + W)

    def act(r, n):
        a = r.uniform(-1, 1, (n, A))
        m = r.uniform(0, 1, (n, A)) < 0.4
        a[m] = -np.abs(a[m])
        return a.astype(np.float32)

    ad = CudaAdapter(env)
    assert_bits_equal(ref.reset(), ad.reset(), "reset obs")
    n_done = 0
    for t in range(90):
        a = act(rng, N)
        o_ref, r_ref, d_ref, _ = ref.step(a)
        o, r, d, _ = ad.step(a)
        assert_bits_equal(d_ref, d, f"dones t={t}")
        st = ad.state()
        for key in ("seg", "ptr", "cash"):
            assert_bits_equal(getattr(ref, key), st[key], f"{key} t={t}")
        for key in ("long_sh", "short_sh", "margin"):
            assert_bits_equal(getattr(ref, key).reshape(-1), st[key], f"{key} t={t}")
        assert_bits_equal(r_ref, r, f"rewards t={t}")
        assert_bits_equal(o_ref, o, f"obs t={t}")
        n_done += int(d_ref.sum())
    assert n_done > 0 and int(env.stats()["n_done"].item()) == n_done


# ------------------------------------------------------------------ BASELINE full sizes -------------
def _check_obs_properties(env, obs, ptr_before_reset, seg_before_reset):
    """Size-independent properties of one step's observation: columns 0..3 are exactly the log-return window
    [ptr, ptr+W) of the env's segment (gathered here with torch indexing), column 4 is constant over the window."""
    W, A = env.num_intervals, env.num_assets
    s = env.series
    n = obs.shape[0]
    o = obs.view(n, W, A, 5)
    row0 = s.seg_start[seg_before_reset.long()] + ptr_before_reset.long()
    table = s.logret.view(s.num_rows, A, 4)
    steps = torch.arange(W, device=obs.device)[None, :]
    chunk = max(1, (1 << 26) // (W * A * 4))                         # EVERY env, 64 Mi values (256 MB f32) at a time
    for lo in range(0, n, chunk):
        rows = row0[lo:lo + chunk, None] + steps
        assert torch.equal(o[lo:lo + chunk, ..., :4], table[rows]), f"window of an env in [{lo}, {lo + chunk})"
    assert torch.equal(o[..., 4], o[:, :1, :, 4].expand(-1, W, -1))


def _assert_full_obs_equals_oracle(obs, o_ref):
    """The WHOLE observation of one step, bit for bit (windows and position-feature values of every env)."""
    got = obs.cpu().numpy().reshape(-1)
    want = np.asarray(o_ref).reshape(-1)
    assert got.dtype == want.dtype and got.shape == want.shape
    assert np.array_equal(got.view(np.uint32 if got.itemsize == 4 else np.uint64), want.view(np.uint32 if want.itemsize == 4 else np.uint64))


def test_config2_full_size_1M_envs():
    """BASELINE config 2 at full size: 1 Mi envs, W=60.  Three steps of the FULL population against the oracle
    (state, rewards, dones exact; the whole 1.26 GB observation bit for bit on the first and the last of them, every env's
    window through the gather property on all three) and a 65 536-env slice in lock-step for 2x252 steps (every env
    auto-resets at least twice).  The oracle is fed the log-return table the GPU staged (series.logret64): the table itself
    is pinned to the reference by test_log_returns_kernel_vs_reference_values."""
    from oracle import oracle as orc
    from finenvs_b200.data import loader

    W, N = 60, 1 << 20
    prices, seg_start, seg_len = _c1_series(W)
    series = loader.stage_series(prices, seg_start, seg_len, W, "cuda:0", torch.float32, keep_logret64=True)
    fs = orc.series_from_prices(prices, seg_start, seg_len, W, logret=series.logret64.cpu().numpy())
    env = _env(series, num_envs=N, seed=42, random_reset="all", random_offset=True, track_stats=True)
    assert env.kernel_name() == "fe_gather_kernel<float>"               # what bench.py times on this workload
    ref = orc.OracleEnv(fs, num_envs=N, seed=42, reset_mode=2, random_offset=True, out_f64=False)
    g = torch.Generator().manual_seed(0)
    total_done = 0
    for t in range(3):
        a = torch.rand((N, 1), generator=g) * 2 - 1
        ptr0, seg0 = env._ptr.clone() + 1, env._seg.clone()
        obs, r, d, _ = env.step(a.cuda())
        o_ref, r_ref, d_ref, _ = ref.step(a.numpy(), want_obs=t != 1)
        assert np.array_equal(d.cpu().numpy(), d_ref) and np.array_equal(r.cpu().numpy(), r_ref)
        for key, t_gpu in (("seg", env._seg), ("ptr", env._ptr), ("cash", env._cash), ("long_sh", env._long),
                           ("short_sh", env._short), ("margin", env._margin)):
            assert np.array_equal(t_gpu.cpu().numpy(), getattr(ref, key)), key
        _check_obs_properties(env, obs, ptr0, seg0)
        if t != 1:
            _assert_full_obs_equals_oracle(obs, o_ref)
        del o_ref
        assert int((env._long * env._short).abs().sum()) == 0          # never long and short at once (:311-314)
        assert bool((env._ptr + W < series.seg_len[env._seg.long()]).all())  # a bar always remains to step onto
        total_done += int(d_ref.sum())
    assert int(env.stats()["n_done"].item()) == total_done
    # slice in lock-step: global ids [base, base+65536) of the same population (shard-invariant draws)
    base, n = 300_000, 65536
    sl = _env(series, num_envs=n, env_id_base=base, total_envs=N, seed=42, random_reset="all", random_offset=True)
    ref_sl = orc.OracleEnv(fs, num_envs=n, env_id_base=base, total_envs=N, seed=42, reset_mode=2, random_offset=True,
                           out_f64=False)
    n_done = _lockstep(sl, ref_sl, 2 * 252, np.random.default_rng(3), obs_every=101)
    assert n_done >= 2 * n
    # the same slice through the round-1 kernel (what `auto` falls back to when 5*W*itemsize is not a multiple of 16)
    pipe = _env(series, num_envs=n, env_id_base=base, total_envs=N, seed=42, random_reset="all", random_offset=True, variant="pipe")
    ref_pipe = orc.OracleEnv(fs, num_envs=n, env_id_base=base, total_envs=N, seed=42, reset_mode=2, random_offset=True,
                             out_f64=False)
    assert pipe.kernel_name() == "fe_pipe_kernel<float,cached>"
    _lockstep(pipe, ref_pipe, 60, np.random.default_rng(4), obs_every=7)


def test_config4_minute_bars_8M_population_shard():
    """BASELINE config 4: 10 M-row minute-bar series, 8 Mi envs with random start offsets, env-sharded over 8 GPUs.
    This GPU holds shard 5 (1 Mi envs); it must equal the oracle run of the same global ids."""
    from oracle import oracle as orc
    from finenvs_b200.data import loader

    W, T, bars = 60, 10_000_000, 390
    rng = np.random.default_rng(20260101)
    prices = np.round(gbm_ohlc(rng, T, 0.0005), 4)
    seg_start, seg_len = loader.regular_segments(T, bars, W)
    series = loader.stage_series(prices, seg_start, seg_len, W, "cuda:0", torch.float32, keep_logret64=True)
    assert series.num_segments == 25640 and series.nbytes() > 126e6    # larger than L2
    fs = orc.series_from_prices(prices, seg_start, seg_len, W, logret=series.logret64.cpu().numpy())
    total, n, rank = 8 << 20, 1 << 20, 5
    env = _env(series, num_envs=n, env_id_base=rank * n, total_envs=total, seed=7, random_reset="all", random_offset=True)
    ref = orc.OracleEnv(fs, num_envs=n, env_id_base=rank * n, total_envs=total, seed=7, reset_mode=2, random_offset=True,
                        out_f64=False)
    assert np.array_equal(env._seg.cpu().numpy(), ref.seg) and np.array_equal(env._ptr.cpu().numpy(), ref.ptr)
    assert len(np.unique(ref.ptr)) == bars                              # every start offset is drawn
    g = torch.Generator().manual_seed(1)
    for t in range(4):
        a = torch.rand((n, 1), generator=g) * 2 - 1
        ptr0, seg0 = env._ptr.clone() + 1, env._seg.clone()
        obs, r, d, _ = env.step(a.cuda())
        o_ref, r_ref, d_ref, _ = ref.step(a.numpy(), want_obs=t == 3)
        assert np.array_equal(d.cpu().numpy(), d_ref) and np.array_equal(r.cpu().numpy(), r_ref)
        assert np.array_equal(env._seg.cpu().numpy(), ref.seg) and np.array_equal(env._ptr.cpu().numpy(), ref.ptr)
        assert np.array_equal(env._cash.cpu().numpy(), ref.cash)
        _check_obs_properties(env, obs, ptr0, seg0)
        if t == 3:   # the whole observation of the last step (after some envs redrew segment and offset), bit for bit
            _assert_full_obs_equals_oracle(obs, o_ref)
        del o_ref
        assert d_ref.sum() > 0                                          # offsets spread the episode ends


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_config4_stream_kernel_long_lockstep(dtype):
    """the kernel that runs BASELINE config 4 (pipe, "stream" flavour: series in HBM) in lock-step with
    the oracle on a 65 536-env slice of the 8 Mi population for 2 x 390 + 10 steps — every env goes through at least two
    auto-resets with (segment, offset) redraws — in both output dtypes."""
    from oracle import oracle as orc
    from finenvs_b200.data import loader

    W, T, bars = 60, 10_000_000, 390
    rng = np.random.default_rng(20260101)
    prices = np.round(gbm_ohlc(rng, T, 0.0005), 4)
    seg_start, seg_len = loader.regular_segments(T, bars, W)
    series = loader.stage_series(prices, seg_start, seg_len, W, "cuda:0", dtype, keep_logret64=True)
    fs = orc.series_from_prices(prices, seg_start, seg_len, W, logret=series.logret64.cpu().numpy())
    total, n, base = 8 << 20, 65536, (3 << 20) + 12345
    kw = dict(num_envs=n, env_id_base=base, total_envs=total, seed=7)
    env = _env(series, random_reset="all", random_offset=True, obs_dtype=dtype, **kw)
    assert env.kernel_name() == f"fe_pipe_kernel<{'double' if dtype == torch.float64 else 'float'},stream>"
    ref = orc.OracleEnv(fs, reset_mode=2, random_offset=True, out_f64=dtype == torch.float64, **kw)
    n_done = _lockstep(env, ref, 2 * bars + 10, np.random.default_rng(5), obs_every=97)
    assert n_done >= 2 * n


@pytest.mark.parametrize("dtype,W,N", [(torch.float64, 60, 40003), (torch.float32, 128, 40001), (torch.float64, 100, 40002),
                                       (torch.float32, 104, 40003)])
def test_gather_kernel_two_part_windows_lockstep(dtype, W, N):
    """Windows longer than one TMA box row (2048 B) — 60 rows of f64, the reference's own observation dtype, or 128 rows of
    f32 — are fetched in two parts, two envs per gather4 (fe_gather_kernel<.., 2>): lock-step with the oracle through more
    than two episodes per env, ragged last tile (N % 32 in 1 .. 3: units with fewer than two live envs), statistics on."""
    from oracle import oracle as orc
    from finenvs_b200.data import loader

    prices, seg_start, seg_len = _c1_series(W, days=60, bars=31, sigma=0.05, seed=W + N)
    series = loader.stage_series(prices, seg_start, seg_len, W, "cuda:0", dtype, keep_logret64=True)
    fs = orc.series_from_prices(prices, seg_start, seg_len, W, logret=series.logret64.cpu().numpy())
    f64 = dtype == torch.float64
    env = _env(series, num_envs=N, seed=9, random_reset="all", random_offset=True, obs_dtype=dtype, track_stats=True)
    assert env.kernel_name() == f"fe_gather_kernel<{'double' if f64 else 'float'}>"
    assert (8 if f64 else 4) * 5 * W > 2048                                # really the two-part path
    ref = orc.OracleEnv(fs, num_envs=N, seed=9, reset_mode=2, random_offset=True, out_f64=f64)
    n_done = _lockstep(env, ref, 70, np.random.default_rng(N), obs_every=5)
    assert n_done >= 2 * N and int(env.stats()["n_done"].item()) == n_done


def test_auto_picks_the_kernels_documented_in_design():
    """`auto`: gather for large populations over an L2-resident series when the window is one or two legal
    TMA rows, pipe otherwise (odd windows, long series), tile for small populations, portfolio for A > 1."""
    from finenvs_b200.data import loader

    def name(W, N, dtype=torch.float32, rows=80 * 37, **kw):
        prices, seg_start, seg_len = _c1_series(W, days=rows // 37, bars=37, sigma=0.05, seed=W)
        series = loader.stage_series(prices, seg_start, seg_len, W, "cuda:0", dtype)
        return _env(series, num_envs=N, obs_dtype=dtype, **kw).kernel_name()

    assert name(60, 40000) == "fe_gather_kernel<float>"
    assert name(60, 40000, variant="pipe") == "fe_pipe_kernel<float,cached>"
    assert name(61, 40000) == "fe_pipe_kernel<float,cached>"            # 20 * 61 is not a multiple of 16
    assert name(50, 40000, torch.float64) == "fe_gather_kernel<double>"
    assert name(60, 40000, torch.float64) == "fe_gather_kernel<double>"        # 40 * 60 = 2400 B: two TMA rows of 15 pitches
    assert name(54, 40000, torch.float64) == "fe_pipe_kernel<double,cached>"   # 2160 B: half a window is not a whole pitch
    assert name(128, 40000) == "fe_gather_kernel<float>"                       # 2560 B in two parts
    assert name(136, 40000) == "fe_pipe_kernel<float,cached>"
    assert name(60, 3000) == "fe_tile_kernel<float>"
    assert name(16, 40000) == "fe_tile_kernel<float>"                   # few rows per env: thread-per-env bookkeeping wins
    assert name(60, 40000, rows=1_000_000 // 37 * 37) == "fe_pipe_kernel<float,cached>"   # table would not stay in L2


def test_config3_full_size_portfolio():
    """BASELINE config 3 at full size: 30 assets, 65 536 envs, W=128 (obs is 5 GB per step): state, rewards, dones exact on
    three steps, every env's windows through the gather property, and the whole observation of the third step bit for
    bit against the multi-asset oracle (itself cross-checked by tests/test_portfolio_restatement.py)."""
    from oracle import oracle as orc
    from finenvs_b200.data import loader

    A, W, N = 30, 128, 65536
    prices, seg_start, seg_len = _portfolio_series(A, W, days=40, bars=252, sigma=0.01, seed=33)
    series = loader.stage_series(prices, seg_start, seg_len, W, "cuda:0", torch.float32, keep_logret64=True)
    fs = orc.series_from_prices(prices, seg_start, seg_len, W, logret=series.logret64.cpu().numpy())
    env = _env(series, num_envs=N, seed=8, random_reset="all", random_offset=True)
    ref = orc.OracleEnv(fs, num_envs=N, seed=8, reset_mode=2, random_offset=True, out_f64=False)
    g = torch.Generator().manual_seed(2)
    for t in range(3):
        a = torch.rand((N, A), generator=g) * 2 - 1
        ptr0, seg0 = env._ptr.clone() + 1, env._seg.clone()
        obs, r, d, _ = env.step(a.cuda())
        o_ref, r_ref, d_ref, _ = ref.step(a.numpy(), want_obs=t == 2)
        assert np.array_equal(d.cpu().numpy(), d_ref) and np.array_equal(r.cpu().numpy(), r_ref)
        assert np.array_equal(env._cash.cpu().numpy(), ref.cash)
        assert np.array_equal(env._margin.cpu().numpy(), ref.margin.reshape(-1))
        assert np.array_equal(env._long.cpu().numpy(), ref.long_sh.reshape(-1))
        _check_obs_properties(env, obs, ptr0, seg0)
        if t == 2:   # all 5 GB of the last step's observation, bit for bit
            _assert_full_obs_equals_oracle(obs, o_ref)
        del obs, o_ref


def test_flat_obs_and_es_env_args():
    """The ES agent needs 2-D fp32 observations, num_eval_envs in the env args and reset_all()
    (evo_agent.py:49-65, parallel_mlp.py:98-103, ES_MLP_Isaac_Gym.py:38)."""
    from finenvs_b200.data import loader

    W = 16
    prices, seg_start, seg_len = _c1_series(W, days=64, bars=40, sigma=0.03)
    series = loader.stage_series(prices, seg_start, seg_len, W, "cuda:0", torch.float32)
    a = _env(series, num_envs=500, seed=1, random_reset="all")
    b = _env(series, num_envs=500, seed=1, random_reset="all", flat_obs=True, num_eval_envs=12)
    args = b.get_env_args()
    assert args["num_observations"] == W * 5 and args["num_eval_envs"] == 12 and "num_eval_envs" not in a.get_env_args()
    act = torch.rand((500, 1), device="cuda") * 2 - 1
    oa, _, _, _ = a.step(act)
    ob, _, _, _ = b.step(act)
    assert ob.shape == (500, W * 5) and ob.dtype == torch.float32 and torch.equal(oa.view(500, -1), ob)
    o0 = b.reset_all()
    assert o0.shape == (500, W * 5) and int(b._ptr.abs().sum()) == 0 and float(b._cash.min()) == 10000.0


@pytest.mark.parametrize("variant", ["pipe", "gather", "split"])
@pytest.mark.parametrize("W,N,dtype", [(60, 1024, torch.float32), (60, 5003, torch.float32), (60, 4099, torch.float64),
                                       (8, 70001, torch.float32), (128, 3000, torch.float32), (3, 999, torch.float32),
                                       (100, 20011, torch.float32), (50, 9001, torch.float64), (4, 1, torch.float32),
                                       (24, 19001, torch.float32), (12, 33, torch.float64)])
def test_pipe_variant_vs_oracle(W, N, dtype, variant):
    """The persistent warp-specialised pipelines: many tiles per block (N >> 148 * 32), ragged last tile / last unit,
    small and larger windows (32/16/8-env tiles; 2- and 3-slot gather rings), both dtypes, resets with redraws — exact
    against the oracle."""
    from oracle import oracle as orc
    from finenvs_b200.data import loader

    if variant == "gather" and _lib_mod().lib().fe_obs_table_bytes(80 * 37, W, int(dtype == torch.float64)) == 0:
        pytest.skip("this window is not a legal TMA row")
    if variant == "pipe" and _lib_mod().lib().fe_pipe_envs(W, int(dtype == torch.float64), 0) == 0:
        pytest.skip("window too large for the pipe variant's rings")

    prices, seg_start, seg_len = _c1_series(W, days=80, bars=37, sigma=0.05, seed=W * 7 + N)
    series = loader.stage_series(prices, seg_start, seg_len, W, "cuda:0", dtype, keep_logret64=True)
    fs = orc.series_from_prices(prices, seg_start, seg_len, W, logret=series.logret64.cpu().numpy())
    env = _env(series, num_envs=N, seed=4, random_reset="all", random_offset=True, obs_dtype=dtype, variant=variant,
               track_stats=True)
    ref = orc.OracleEnv(fs, num_envs=N, seed=4, reset_mode=2, random_offset=True, out_f64=dtype == torch.float64)
    n_done = _lockstep(env, ref, 60, np.random.default_rng(N), obs_every=4)
    assert n_done > 0 and int(env.stats()["n_done"].item()) == n_done


@pytest.mark.parametrize("N,A,dtype,W", [(70001, 1, torch.float32, 12), (40000, 1, torch.float64, 12), (9000, 1, torch.float32, 12),
                                         (33000, 3, torch.float32, 12), (19007, 1, torch.float32, 12),
                                         (40003, 1, torch.float32, 24),     # gather kernel: dones packed by the bookkeepers
                                         (40003, 1, torch.float32, 26),     # pipe kernel, 32-env tiles: same
                                         (30011, 1, torch.float64, 130)])   # pipe kernel, 4-env tiles: separate pack kernel
def test_step_host_pipeline_equals_device_step(N, A, dtype, W):
    """fe_step_host cuts the envs into chunks on side streams (upload / kernel / download overlapped): same
    results as the one-launch device-resident step, for ragged chunk counts, both dtypes, A > 1, with statistics."""
    from finenvs_b200.data import loader

    if A == 1:
        prices, seg_start, seg_len = _c1_series(W, days=50, bars=30, sigma=0.04, seed=N)
    else:
        prices, seg_start, seg_len = _portfolio_series(A, W, 50, 30, 0.04, N)
    series = loader.stage_series(prices, seg_start, seg_len, W, "cuda:0", dtype)
    kw = dict(num_envs=N, seed=21, random_reset="all", random_offset=True, obs_dtype=dtype, track_stats=True)
    a_env, b_env = _env(series, **kw), _env(series, **kw)
    g = torch.Generator().manual_seed(N)
    for t in range(45):
        a = (torch.rand((N, A), generator=g) * 2 - 1)
        o1, r1, d1, _ = a_env.step(a.cuda())
        packed = t % 3 == 2     # every third step returns the dones bit-packed (fe_step_host_packed)
        o2, r2, d2, _ = b_env.step_host(a.pin_memory() if t % 2 else a, packed_dones=packed)
        assert r2.device.type == "cpu" and d2.device.type == "cpu"
        if packed:
            assert d2.dtype == torch.uint8 and d2.numel() == 4 * ((N + 31) // 32)
            d2 = b_env.unpack_dones(d2)
        assert torch.equal(o1, o2) and torch.equal(r1.cpu(), r2) and torch.equal(d1.cpu(), d2)
        assert torch.equal(b_env._host_bufs[1], r1) and torch.equal(b_env._host_bufs[2], d1)   # device copies too
        for k in ("_seg", "_ptr", "_cash", "_long", "_short", "_margin"):
            assert torch.equal(getattr(a_env, k), getattr(b_env, k)), k
    sa, sb = a_env.stats(), b_env.stats()
    assert int(sa["n_done"]) == int(sb["n_done"]) > 0 and int(sa["sum_len"]) == int(sb["sum_len"])


@pytest.mark.parametrize("variant,A,N", [("tile", 1, 3001), ("pipe", 1, 40000), ("split", 1, 5003), ("gather", 1, 40003), ("auto", 3, 2500)])
@pytest.mark.parametrize("params", [(40, 500.0, 0.37, 2.25, 0.4), (1, 250000.0, 0.0, 1.0, 0.1), (12, 20000.0, 0.05, 1.2, 0.3)])
def test_non_default_parameters_vs_oracle(variant, A, N, params):
    """Every constructor parameter of the reference off its default (time_series_env.py:20-26; the three golden
    trace_par_* files pin the same parameter sets to the real reference): kernels vs oracle at larger N, A = 3 included."""
    from oracle import oracle as orc
    from finenvs_b200.data import loader

    W = 12 if variant == "gather" else 10
    ms, sb, com, imr, mmr = params
    if A == 1:
        prices, seg_start, seg_len = _c1_series(W, days=60, bars=33, sigma=0.06, seed=int(ms) + N)
    else:
        prices, seg_start, seg_len = _portfolio_series(A, W, 60, 33, 0.06, int(ms) + N)
    series = loader.stage_series(prices, seg_start, seg_len, W, "cuda:0", torch.float32, keep_logret64=True)
    fs = orc.series_from_prices(prices, seg_start, seg_len, W, logret=series.logret64.cpu().numpy())
    env = _env(series, num_envs=N, seed=6, random_reset="all", random_offset=True, variant=variant, track_stats=True,
               max_shares=int(ms), starting_balance=sb, per_share_commission=com, initial_margin_requirement=imr,
               maintenance_margin_requirement=mmr)
    ref = orc.OracleEnv(fs, num_envs=N, seed=6, reset_mode=2, random_offset=True, out_f64=False, max_shares=int(ms),
                        starting_balance=sb, commission=com, imr=imr, mmr=mmr)
    rng = np.random.default_rng(N)
    ad = CudaAdapter(env)
    assert_bits_equal(ref.reset(), ad.reset(), "reset obs")
    n_done = 0
    for t in range(70):
        a = np.clip(rng.normal(-0.2, 0.8, (N, A)), -1.3, 1.3).astype(np.float32)
        o_ref, r_ref, d_ref, _ = ref.step(a if A > 1 else a.reshape(-1))
        o, r, d, _ = ad.step(a)
        assert_bits_equal(d_ref, d, f"dones t={t}")
        st = ad.state()
        for key in ("seg", "ptr", "cash"):
            assert_bits_equal(getattr(ref, key), st[key], f"{key} t={t}")
        for key in ("long_sh", "short_sh", "margin"):
            assert_bits_equal(getattr(ref, key).reshape(-1), st[key], f"{key} t={t}")
        assert_bits_equal(r_ref, r, f"rewards t={t}")
        if t % 5 == 0:
            assert_bits_equal(o_ref, o, f"obs t={t}")
        n_done += int(d_ref.sum())
    assert n_done > 0 and int(env.stats()["n_done"].item()) == n_done


@pytest.mark.parametrize("trace,instr,how", [("trace_ibm_w60.npz", "IBM", "env_var"), ("trace_oih_w60_eval.npz", "OIH", "path"),
                                             ("trace_spy_w390.npz", "SPY", "path"), ("kat_ibm_w390.npz", "IBM", "env_var")])
def test_drop_in_constructor_from_csv_replays_the_reference(tmp_path, monkeypatch, trace, instr, how):
    """The constructor a reference user calls — TimeSeriesEnv("IBM", "dummy", num_intervals=60) on a data directory
    with the reference's CSV (:15-45, :80-216) — end to end: CSV loader, staging, GPU log-returns, reference defaults
    (num_envs = days (+1 when training), last-env redraws).  State, rewards and dones replay the reference's trace
    exactly; observations to 1e-12 absolute in f64 (fe_log_returns' log vs torch's: <= 2 ulp of a value of O(1))."""
    from finenvs_b200.environments import TimeSeriesEnv

    z = load_trace(trace)
    zc = load_trace("dummy_csv.npz")
    W = int(z["window"])
    base = tmp_path / "root"
    d = base / instr if how == "env_var" else tmp_path / f"data_{instr}"      # a path containing "data" is used verbatim
    d.mkdir(parents=True)
    with open(d / "SYN_dummy_2020.csv", "w") as f:                            # any *dummy*.csv (:60-73)
        for date, time, ohlc, vol in zip(zc[f"{instr}_csv_date"], zc[f"{instr}_csv_time"], zc[f"{instr}_csv_ohlc"],
                                         zc[f"{instr}_csv_volume"]):
            f.write(f"{date.decode()},{time.decode()},{float(ohlc[0])!r},{float(ohlc[1])!r},{float(ohlc[2])!r},{float(ohlc[3])!r},{vol}\n")
    if how == "env_var":
        monkeypatch.setenv("FINENVS_DATA_DIR", str(base))
        name = instr
    else:
        name = str(d)
    env = TimeSeriesEnv(name, "dummy", W, evaluate=bool(z["evaluate"]), device_id=0, obs_dtype=torch.float64, seed=int(z["seed"]))
    N = len(z["seg_init"])
    assert env.num_envs == N and env.get_env_args() == {"env_name": name, "num_envs": N, "num_observations": 5,
                                                       "num_actions": 1, "sequence_length": W}
    assert np.array_equal(env.env_indices.cpu().numpy(), z["seg_init"])       # incl. the drawn day of the extra env (:253)
    obs = env.reset()
    assert obs.shape == (N, W, 5) and obs.dtype == torch.float64
    np.testing.assert_allclose(obs.cpu().numpy(), z["obs_reset"], rtol=0, atol=1e-12)
    obs_steps = {int(t): k for k, t in enumerate(z["obs_steps"])}
    for t in range(z["actions"].shape[0]):
        o, r, dn, info = env.step(torch.from_numpy(z["actions"][t]).view(N, 1).cuda())
        assert np.array_equal(dn.cpu().numpy(), z["dones"][t]), t
        assert np.array_equal(r.cpu().numpy(), z["rewards"][t]), t
        st = z["states"][t]
        assert np.array_equal(env.env_indices.cpu().numpy(), st[0].astype(np.int64)), t
        assert np.array_equal(env.env_pointers.cpu().numpy(), st[1].astype(np.int64)), t
        assert np.array_equal(env.cash.view(-1).cpu().numpy(), st[2].astype(np.float32)), t
        assert np.array_equal(env.margin.view(-1).cpu().numpy(), st[5]), t
        if t in obs_steps:
            np.testing.assert_allclose(o.cpu().numpy(), z["obs"][obs_steps[t]], rtol=0, atol=1e-12)
        if "info_steps" in z and t in set(int(x) for x in z["info_steps"]):
            k = list(z["info_steps"]).index(t)
            assert np.array_equal(info["returns"].cpu().numpy(), z["info_returns"][k])


@pytest.mark.parametrize("dtype,rows", [(torch.float32, 3_300_000), (torch.float64, 1_700_000)])
def test_pipe_stream_flavour_vs_oracle(dtype, rows):
    """A log-return table beyond 48 MB switches the pipe variant to its "stream" flavour (cp.async in-ring): both
    output dtypes against the oracle, ragged last tile included."""
    from oracle import oracle as orc
    from finenvs_b200.data import loader

    W, N, bars = 60, 40003, 390
    rng = np.random.default_rng(rows)
    prices = np.round(gbm_ohlc(rng, rows, 0.002), 4)
    seg_start, seg_len = loader.regular_segments(rows, bars, W)
    series = loader.stage_series(prices, seg_start, seg_len, W, "cuda:0", dtype, keep_logret64=True)
    env = _env(series, num_envs=N, seed=8, random_reset="all", random_offset=True, obs_dtype=dtype, variant="pipe")
    assert env.kernel_name().endswith("stream>"), env.kernel_name()
    fs = orc.series_from_prices(prices, seg_start, seg_len, W, logret=series.logret64.cpu().numpy())
    ref = orc.OracleEnv(fs, num_envs=N, seed=8, reset_mode=2, random_offset=True, out_f64=dtype == torch.float64)
    _lockstep(env, ref, 12, np.random.default_rng(1), obs_every=3)
