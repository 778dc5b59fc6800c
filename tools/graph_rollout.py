#!/usr/bin/env python
"""Eager vs CUDA-graph-replayed rollouts for small populations (BASELINE config 1: 1024 envs, W=60).

    python tools/graph_rollout.py [--envs 1024] [--window 60] [--steps 64] [--replays 50]

At this size one step is ~5 us of GPU work, so the loop is bound by Python + launch overhead; CapturedRollout replays
`steps` x (policy -> step) with one launch.  Prints one JSON line: env-steps/s for both, same policy, same results
(the final observation / state of the two envs are compared bit for bit)."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1024)
    ap.add_argument("--window", type=int, default=60)
    ap.add_argument("--steps", type=int, default=64, help="steps per captured graph")
    ap.add_argument("--replays", type=int, default=50)
    args = ap.parse_args()

    import bench
    from finenvs_b200.data import loader
    from finenvs_b200.environments import TimeSeriesEnv

    W, N, K = args.window, args.envs, args.steps
    prices, seg_start, seg_len, _ = bench.make_series("c2", W)
    series = loader.stage_series(prices, seg_start, seg_len, W, "cuda:0", torch.float32)
    kw = dict(num_intervals=W, device_id=0, series=series, num_envs=N, seed=1, random_reset="all", random_offset=True)
    a, b = TimeSeriesEnv("eager", **kw), TimeSeriesEnv("graph", **kw)

    def policy(obs):   # a stand-in policy of three small torch ops
        return torch.tanh(obs[:, -1, 3:4] * 25.0 - obs[:, 0, 4:5])

    roll = b.capture_rollout(policy, K)
    obs = a.reset()
    for _ in range(K):   # warm-up
        obs, _, _, _ = a.step(policy(obs))
    roll.replay()
    torch.cuda.synchronize()

    t0 = time.perf_counter()
    for _ in range(args.replays * K):
        obs, r, d, _ = a.step(policy(obs))
    torch.cuda.synchronize()
    t_eager = time.perf_counter() - t0

    t0 = time.perf_counter()
    for i in range(args.replays):
        o_g, _, _ = roll.replay(refresh_obs=False)
    torch.cuda.synchronize()
    t_graph = time.perf_counter() - t0

    same = bool(torch.equal(o_g, obs) and torch.equal(a._cash, b._cash) and torch.equal(a._ptr, b._ptr))
    n = args.replays * K * N
    print(json.dumps({
        "workload": f"single-asset env, {N} envs, W={W}, policy of 3 torch ops, {K} steps per graph, {args.replays} replays",
        "eager_env_steps_per_sec": n / t_eager, "eager_us_per_step": t_eager / (args.replays * K) * 1e6,
        "graph_env_steps_per_sec": n / t_graph, "graph_us_per_step": t_graph / (args.replays * K) * 1e6,
        "speedup": t_eager / t_graph, "identical_results": same, "kernel": a.kernel_name(),
    }), flush=True)


if __name__ == "__main__":
    main()
