"""CUDA-graph captured rollouts (finenvs_b200.environments.CapturedRollout, fe_step_captured):
a replayed graph must leave the env, and return observations / rewards / dones, exactly as eager stepping does —
including the Philox redraws, whose step ordinal lives in device memory inside the graph."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _series(W, A=1, rows=3000, bars=50, sigma=0.05, seed=2, dtype=torch.float32):
    from finenvs_b200.data import loader
    from parity_utils import gbm_ohlc

    rng = np.random.default_rng(seed)
    if A == 1:
        prices = np.round(gbm_ohlc(rng, rows, sigma), 4)
    else:
        prices = np.stack([np.round(gbm_ohlc(rng, rows, sigma, s0=30.0 + 11 * a), 4) for a in range(A)], axis=1)
    seg_start, seg_len = loader.regular_segments(rows, bars, W)
    return loader.stage_series(prices, seg_start, seg_len, W, "cuda:0", dtype)


def _policy(env):
    A = env.num_acts

    def policy(obs):  # deterministic, depends on the whole current observation
        x = obs.reshape(env.num_envs, env.num_intervals, A, 5)
        return torch.tanh(x[:, -1, :, 3] * 40.0 + x[:, :, :, 0].mean(dim=1) * 10.0 - x[:, 0, :, 4] * 3.0).float()

    return policy


@pytest.mark.parametrize("N,W,A,variant,dtype", [(3001, 12, 1, "auto", torch.float32), (40000, 8, 1, "pipe", torch.float32),
                                                 (2000, 12, 1, "direct", torch.float64), (1500, 8, 3, "auto", torch.float32)])
def test_replay_equals_eager(N, W, A, variant, dtype):
    from finenvs_b200.environments import TimeSeriesEnv

    series = _series(W, A, dtype=dtype)
    kw = dict(num_intervals=W, device_id=0, series=series, num_envs=N, seed=9, random_reset="all", random_offset=True,
              track_stats=True, variant=variant, obs_dtype=dtype)
    a, b = TimeSeriesEnv("eager", **kw), TimeSeriesEnv("graph", **kw)
    K = 7
    roll = b.capture_rollout(_policy(b), K)
    for k in ("_seg", "_ptr", "_cash", "_long", "_short", "_margin"):   # capturing must not move the env
        assert torch.equal(getattr(a, k), getattr(b, k)), k
    pol = _policy(a)
    obs = a.reset()
    n_done = 0
    for rep in range(6):
        refresh = None if rep < 3 else rep == 4          # default: only the first replay refreshes
        if refresh or rep == 0:
            obs = a.reset()                              # refresh_obs == what reset() returns (post-auto-reset state)
        o_g, r_g, d_g = roll.replay(refresh_obs=refresh)
        for t in range(K):
            act = pol(obs)
            obs, r, d, _ = a.step(act)
            assert torch.equal(roll.actions[t].view_as(act), act), (rep, t)
            assert torch.equal(r_g[t], r) and torch.equal(d_g[t], d), (rep, t)
            n_done += int(d.sum())
        assert torch.equal(o_g, obs), rep
        for k in ("_seg", "_ptr", "_cash", "_long", "_short", "_margin", "_ep_return", "_ep_len"):
            assert torch.equal(getattr(a, k), getattr(b, k)), (rep, k)
        assert a.step_count == b.step_count
    assert n_done > 0 and int(b.stats()["n_done"]) == n_done == int(a.stats()["n_done"])


def test_evaluate_policy_equals_the_reference_style_loop():
    """PPO_LSTM_testing_SPY.py:43-52: step until info carries "returns"."""
    from finenvs_b200.environments import TimeSeriesEnv

    W = 6
    series = _series(W, rows=2400, bars=40, sigma=0.08, seed=5)
    kw = dict(num_intervals=W, evaluate=True, device_id=0, series=series)
    a, b = TimeSeriesEnv("loop", **kw), TimeSeriesEnv("graph", **kw)
    assert a.num_envs == series.num_segments
    pol = _policy(a)
    states, steps = a.reset(), 0
    while True:
        states, rewards, dones, info = a.step(pol(states))
        steps += 1
        if "returns" in info:
            break
    out = b.evaluate_policy(_policy(b), steps_per_replay=16)
    assert out["steps"] >= steps and out["steps"] - steps < 16
    assert torch.equal(out["returns"], info["returns"])
    assert float(out["returns"].abs().sum()) > 0
    # the metrics were reset in place: the same captured graph evaluates again (next checkpoint) with equal result
    # once both envs are back on fresh episodes
    a.reset_all(), b.reset_all()
    states = a.reset()
    while True:
        states, rewards, dones, info = a.step(pol(states))
        if "returns" in info:
            break
    out2 = b.evaluate_policy(None, rollout=out["rollout"])
    assert torch.equal(out2["returns"], info["returns"])


def test_eager_evaluation_keeps_the_state_addresses_a_captured_graph_holds():
    """reset_evaluation_metrics() used to REPLACE the episode-return tensor; a CapturedRollout has the old
    device pointer baked into its graph, so after one eagerly completed evaluation its replays accumulated into freed
    memory.  The metrics are zeroed in place now and info["returns"] is a copy."""
    from finenvs_b200.environments import TimeSeriesEnv

    W = 6
    series = _series(W, rows=2400, bars=40, sigma=0.08, seed=5)
    kw = dict(num_intervals=W, evaluate=True, device_id=0, series=series)
    a, b = TimeSeriesEnv("eager", **kw), TimeSeriesEnv("graph", **kw)
    roll = b.capture_rollout(_policy(b), 8)             # graph captured BEFORE any evaluation completes
    ptr_before = b.episode_returns.data_ptr()
    pol_a, pol_b = _policy(a), _policy(b)

    def eager_eval(env, pol):
        states = env.reset()
        while True:
            states, rewards, dones, info = env.step(pol(states))
            if "returns" in info:
                return info["returns"]

    ra, rb = eager_eval(a, pol_a), eager_eval(b, pol_b)  # both complete one evaluation eagerly
    assert torch.equal(ra, rb) and b.episode_returns.data_ptr() == ptr_before
    assert float(b.episode_returns.abs().sum()) == 0.0 and float(rb.abs().sum()) > 0     # zeroed in place, copy handed out
    a.reset_all(), b.reset_all()
    ra2 = eager_eval(a, pol_a)
    out = b.evaluate_policy(None, rollout=roll)          # the old graph must still write where the env reads
    assert torch.equal(out["returns"], ra2)
