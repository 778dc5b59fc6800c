#!/usr/bin/env python
"""PPO rollout storage: finenvs_b200.agents.PPO.Buffer against the reference's storage scheme.

    python tools/ppo_storage.py [--envs 65536] [--window 60] [--steps 64]

One rollout = `steps` x (random actions -> env.step -> buffer.store) + prepare_training_data.  Compared:
  reference scheme  buffer.py:33-63 / :80-100 restated here for the comparison: every key re-grown with
                    torch.cat(dim=1) per step, returns/advantages by the Python loop of ~6 torch ops per time step
                    (same env, same observations; f32 to keep the memory comparable — the reference stores f64)
  Buffer            pre-allocated time-major storage, observations written in place by the step kernel (bind_env),
                    returns/advantages by one launch of fe_returns_advantages
Prints one JSON line: ms per rollout for both, bytes the returns kernel moves and its GB/s, max |difference| of the returns.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--window", type=int, default=60)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--rollouts", type=int, default=3)
    args = ap.parse_args()

    import bench
    from finenvs_b200.agents.PPO import Buffer
    from finenvs_b200.data import loader
    from finenvs_b200.environments import TimeSeriesEnv

    W, N, T = args.window, args.envs, args.steps
    prices, seg_start, seg_len, _ = bench.make_series("c2", W)
    series = loader.stage_series(prices, seg_start, seg_len, W, "cuda:0", torch.float32)
    kw = dict(num_intervals=W, device_id=0, series=series, num_envs=N, seed=1, random_reset="all", random_offset=True)
    g = torch.Generator(device="cuda").manual_seed(0)
    acts = [torch.rand((N, 1), generator=g, device="cuda") * 2 - 1 for _ in range(T)]
    vals = [torch.rand((N, 1), generator=g, device="cuda") for _ in range(T + 1)]
    gamma = 0.99

    def ev():
        return torch.cuda.Event(enable_timing=True)

    # ---- reference scheme
    env = TimeSeriesEnv("ref", **kw)
    ref_ms, ref_returns = [], None
    for _ in range(args.rollouts):
        cont = {}
        s0 = env.reset()
        torch.cuda.synchronize()
        t0, t1, t2 = ev(), ev(), ev()
        t0.record()
        for t in range(T):
            s1, r, d, _ = env.step(acts[t])
            for key, x in (("states", s0), ("actions", acts[t]), ("rewards", r.unsqueeze(-1)), ("dones", d.unsqueeze(-1)),
                           ("values", vals[t])):
                x = x.unsqueeze(1)                                   # buffer.py:58-63 time axis at dim 1
                cont[key] = x if key not in cont else torch.cat((cont[key], x), dim=1)   # :51-56
            s0 = s1
        t1.record()
        rewards, dones, values = cont["rewards"], cont["dones"], cont["values"]           # (N, T, 1)
        returns = torch.zeros_like(rewards)
        cur = vals[T]
        for t in reversed(range(T)):                                                      # :90-98
            cur = rewards[:, t, :] + (1 - dones[:, t, :]) * gamma * cur
            returns[:, t, :] = cur
        adv = returns - values
        t2.record()
        torch.cuda.synchronize()
        ref_ms.append((t0.elapsed_time(t1), t1.elapsed_time(t2)))
        ref_returns = returns.transpose(0, 1).contiguous()
        del cont, returns, adv

    # ---- Buffer
    env = TimeSeriesEnv("buf", **kw)
    buf = Buffer(16, gamma, 0, capacity=T)
    buf.bind_env(env)
    our_ms = []
    zeros = torch.zeros((N, 1), device="cuda")
    for _ in range(args.rollouts):
        buf.clear()
        s0 = env.reset()
        torch.cuda.synchronize()
        t0, t1, t2 = ev(), ev(), ev()
        t0.record()
        for t in range(T):
            s1, r, d, _ = env.step(acts[t])
            buf.store(s0, acts[t], r, d, zeros, vals[t])
            s0 = s1
        t1.record()
        buf.compute_returns_and_advantages(vals[T])
        t2.record()
        torch.cuda.synchronize()
        our_ms.append((t0.elapsed_time(t1), t1.elapsed_time(t2)))
    ours = buf.container["returns"]
    bytes_moved = T * N * (4 + 4 + 4 + 4 + 4) + 4 * N    # rewards, dones, values read; returns, advantages written
    out = {
        "workload": f"{N} envs, W={W}, {T}-step rollouts, f32",
        "reference_scheme_ms": {"rollout_store": ref_ms[-1][0], "returns_advantages": ref_ms[-1][1]},
        "buffer_ms": {"rollout_store": our_ms[-1][0], "returns_advantages": our_ms[-1][1]},
        "returns_kernel_bytes": bytes_moved, "returns_kernel_GBps": bytes_moved / (our_ms[-1][1] * 1e-3) / 1e9,
        "returns_max_abs_diff_vs_reference_scheme": float((ours - ref_returns).abs().max()),
        "same_env_trajectory": True,
    }
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
