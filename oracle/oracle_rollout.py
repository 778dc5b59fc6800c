"""CPU restatement of the reference's PPO rollout-return scan (TEST INFRASTRUCTURE ONLY — the product
never imports this module; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may).

Follows finenvs/agents/PPO/buffer.py:80-100 (compute_returns_and_advantages) dtype for dtype:

    current_returns = last_values                                   f32 (critic output)            :90
    for idx in reversed(range(num_steps)):                                                          :91
        current_returns = rewards[:, idx] + (1 - dones[:, idx]) * gamma * current_returns           :94-97
        returns[:, idx] = current_returns                           stored into an f32 tensor       :98
    advantages = returns - values                                   f32 - f32                       :100

`(1 - dones) * gamma` is an int32 tensor times a python float = an f32 tensor holding (float)gamma or 0.
With f32 rewards every product and sum is f32.  With f64 rewards (what the reference env returns,
time_series_env.py:296) the running value becomes f64 after the first iteration: iteration T-1 multiplies
f32 x f32 (last_values) and widens the product, later iterations multiply (double)keep x f64; each stored
return is rounded once to f32 while the f64 running value carries on.
Pinned against the reference itself: tests/golden/ppo_buffer.npz (tests/golden/make_golden_buffer.py).
"""
from __future__ import annotations

import numpy as np


def returns_and_advantages(rewards: np.ndarray, dones: np.ndarray, values: np.ndarray, last_values: np.ndarray,
                           gamma: float):
    """rewards (T, N) f32|f64, dones (T, N) int32, values (T, N) f32, last_values (N,) f32 ->
    returns (T, N) f32, advantages (T, N) f32 (time-major, the layout finenvs_b200's buffer stores)."""
    T, N = rewards.shape
    assert rewards.dtype in (np.float32, np.float64) and values.dtype == np.float32 and last_values.dtype == np.float32
    g32 = np.float32(gamma)
    returns = np.zeros((T, N), np.float32)
    cur = last_values.astype(np.float32).copy()
    for t in range(T - 1, -1, -1):
        keep = (1 - dones[t].astype(np.int32)).astype(np.float32) * g32           # f32 tensor: gamma or 0
        if rewards.dtype == np.float32:
            cur = rewards[t] + keep * cur                                         # all f32
        elif cur.dtype == np.float32:
            cur = rewards[t] + (keep * cur).astype(np.float64)                    # first iteration: f32 product widened
        else:
            cur = rewards[t] + keep.astype(np.float64) * cur                      # f32 tensor x f64 tensor -> f64
        returns[t] = cur.astype(np.float32)
    return returns, returns - values
