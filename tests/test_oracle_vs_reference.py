"""Live differential run: the REAL reference (imported from the read-only checkout, CPU) against the
oracle, stepped in lock-step.  Skipped where the checkout does not exist (the GPU box); the golden
traces carry the same evidence there."""
import numpy as np
import pytest

from oracle import oracle as orc
from oracle import ref_harness as rh
from parity_utils import assert_bits_equal, assert_state_equal, day_labels, gbm_ohlc, oracle_state

pytestmark = pytest.mark.skipif(not rh.available(), reason="reference checkout not present")


def _lockstep(csv, W, steps, seed, evaluate=False, widen=None, action_fn=None, ref_kw=None, orc_kw=None):
    r = rh.RefEnv(csv, "dummy", W, seed=seed, evaluate=evaluate, **(ref_kw or {}))
    try:
        fs_ref = rh.flat_series_from_ref(r)
        fs = orc.load_csv(csv, W)  # the oracle's own loader restatement
        assert np.array_equal(fs.prices, fs_ref.prices)
        assert np.array_equal(fs.seg_start, fs_ref.seg_start) and np.array_equal(fs.seg_len_raw, fs_ref.seg_len_raw)
        np.testing.assert_allclose(fs.logret, fs_ref.logret, rtol=1e-15, atol=1e-15)
        pe, le = fs_ref.padded()
        assert np.array_equal(pe, r.env.price_environments.numpy(), equal_nan=True)
        assert np.array_equal(le, r.env.log_return_environments.numpy(), equal_nan=True)
        seg_init = None
        if widen:
            seg_init = np.arange(widen) % fs.num_segments
            r.widen(seg_init)
        o = orc.OracleEnv(fs_ref, num_envs=widen, evaluate=evaluate, seed=seed, seg_init=seg_init, **(orc_kw or {}))
        assert_state_equal(r.state(), oracle_state(o), "init")
        assert_bits_equal(r.reset(), o.reset(), "reset obs")
        rng = np.random.default_rng(seed + 1)
        n_done = 0
        for t in range(steps):
            a = action_fn(rng, o.N) if action_fn else rng.uniform(-1, 1, o.N).astype(np.float32)
            ro, rr, rd, ri = r.step(a)
            oo, orr, od, oi = o.step(a)
            assert_bits_equal(rd, od, f"dones t={t}")
            assert_bits_equal(rr, orr, f"rewards t={t}")
            assert_bits_equal(ro, oo, f"obs t={t}")
            assert_state_equal(r.state(), oracle_state(o), f"t={t}")
            assert set(ri) == set(oi)
            if ri:
                assert_bits_equal(ri["returns"], oi["returns"], "returns")
            n_done += int(od.sum())
        return n_done
    finally:
        r.close()


@pytest.mark.parametrize("name,W", [("IBM", 390), ("OIH", 60), ("SPY", 4)])
def test_dummy_csvs_training_mode(name, W):
    assert _lockstep(rh.data_dir(name) + "/dummy.csv", W, 420, seed=21) > 0


def test_dummy_csv_evaluate_mode():
    _lockstep(rh.data_dir("OIH") + "/dummy.csv", 60, 300, seed=22, evaluate=True)


def test_widened_adversarial_series(tmp_path):
    rng = np.random.default_rng(99)
    bars = [30, 7, 30, 22, 1, 30, 30, 16]
    dates, times = day_labels(len(bars), bars)
    p = str(tmp_path / "adv.csv")
    rh.write_csv(p, dates, times, gbm_ohlc(rng, sum(bars), 0.12))

    def act(r, n):
        a = r.uniform(-1, 1, n)
        m = r.uniform(0, 1, n) < 0.7
        a[m] = -np.abs(a[m])
        return a.astype(np.float32)

    assert _lockstep(p, 6, 250, seed=23, widen=64, action_fn=act) > 500


@pytest.mark.parametrize("case", range(40))
def test_fuzzed_series_and_parameters(tmp_path, case):
    """Seeded fuzz: ragged days (1-bar and short days included), volatility from calm to violent, every constructor
    parameter of the reference moved off its default (time_series_env.py:15-29), both modes — bit for bit."""
    rng = np.random.default_rng(1000 + case)
    W = int(rng.integers(1, 13))
    days = int(rng.integers(5, 12))
    bars = [int(b) for b in rng.choice([1, 2, 3, 5, 9, 14, 20, 31], size=days)]
    bars[0] = max(bars[0], W + 1)                       # the first day only provides history (:134)
    sigma = float(rng.choice([0.004, 0.02, 0.08, 0.2]))
    dates, times = day_labels(days, bars)
    p = str(tmp_path / f"fuzz{case}.csv")
    rh.write_csv(p, dates, times, gbm_ohlc(rng, sum(bars), sigma, s0=float(rng.choice([3.0, 40.0, 900.0]))))
    max_shares = int(rng.choice([1, 5, 40]))
    balance = float(rng.choice([500.0, 10000.0, 250000.0]))
    commission = float(rng.choice([0.0, 0.01, 0.37]))
    imr = float(rng.choice([1.0, 1.5, 2.25]))
    mmr = float(rng.choice([0.1, 0.25, 0.4]))
    evaluate = bool(case % 3 == 0)
    ref_kw = dict(max_shares=max_shares, starting_balance=balance, per_share_commission=commission,
                  initial_margin_requirement=imr, maintenance_margin_requirement=mmr)
    orc_kw = dict(max_shares=max_shares, starting_balance=balance, commission=commission, imr=imr, mmr=mmr)
    bias = float(rng.uniform(-0.6, 0.6))

    def act(r, n):                                      # exceeds [-1, 1] now and then: the env clamps (:301)
        return np.clip(r.normal(bias, 0.8, n), -1.3, 1.3).astype(np.float32)

    _lockstep(p, W, 160, seed=2000 + case, evaluate=evaluate, widen=48, action_fn=act, ref_kw=ref_kw, orc_kw=orc_kw)
