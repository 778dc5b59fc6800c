#!/usr/bin/env python
"""BASELINE config 5: ES population rollout on the env-sharded trading env, fitness reduced over NCCL.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P \
        tools/es_rollout.py --envs-per-gpu 524288 --steps 64 --generations 3

The reference's loop (examples/isaac_gym/ES_MLP_Isaac_Gym.py:30-38) on finenvs_b200's drop-ins:
`EvoAgent.step` -> `TimeSeriesEnv.step_lazy` -> `EvoAgent.store_async`, then `EvoAgent.train()` per generation.
Per step and GPU that is three kernel launches and no host synchronisation:

  fe_es_forward   perturbed-MLP policy for every env, reading each env's window straight from the staged series
                  (lazy observation handles: the (N, W*5) observation tensor is never materialised) and the pair's
                  fp16 perturbation from HBM
  fe_lazy_kernel  the env step (state transition, reward, done, auto-reset) -> 12-byte observation handle per env
  fe_es_store     running returns + append of finished episodes to the device list

`train()`: all-gather of the finished-episode lists over NVLink -> ONE global centred-rank transform (identical on every
rank) -> fe_es_gradient over this rank's perturbations -> all-reduce(sum) of the parameter-sized gradient -> Adam.
Each rank holds a self-contained [positive | negative | eval] block of the population (pairs never straddle GPUs).

`--mode dense` runs the same loop through materialised observations (env.step -> (N, W*5) tensor -> forward) for
comparison.  Prints one JSON line from rank 0.
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
import torch.distributed as dist  # noqa: E402


def run(rank, world, local_rank, envs_per_gpu=524288, eval_envs_per_gpu=0, steps=64, generations=3, window=60, hidden=(8,),
        workload="c4", mode="lazy", sigma=0.05, lr=0.01, make_series=None):
    """The rollout + train loop on this rank's GPU (process group already initialised when world > 1); returns the record
    (every rank gets the same per-generation numbers: times are the max over ranks).  bench.py calls this for its
    `also.c5` block."""
    if make_series is None:
        import bench
        make_series = bench.make_series
    from finenvs_b200 import parallel as par
    from finenvs_b200.agents.ES import EvoAgent
    from finenvs_b200.data import loader
    from finenvs_b200.environments import TimeSeriesEnv

    dev = torch.device("cuda", local_rank)
    W, n, n_eval = window, envs_per_gpu, eval_envs_per_gpu
    assert (n - n_eval) % 2 == 0, "mirrored sampling needs an even training population per GPU"
    total = n * world
    pairs = (n - n_eval) // 2
    prices, seg_start, seg_len, _ = make_series(workload, W)
    series = loader.stage_series(prices, seg_start, seg_len, W, str(dev), torch.float32)
    del prices
    env = TimeSeriesEnv("es", num_intervals=W, device_id=local_rank, series=series, num_envs=n, env_id_base=rank * n,
                        total_envs=total, seed=5, random_reset="all", random_offset=True, flat_obs=True,
                        num_eval_envs=n_eval, track_stats=True)
    torch.manual_seed(3)   # same initial parameters on every rank
    agent = EvoAgent(env.get_env_args(), hidden_dims=tuple(hidden), learning_rate=lr, noise_std_dev=sigma,
                     write_to_csv=False, device_id=local_rank, seed=11, env_id_base=rank * n, total_envs=total,
                     pair_id_base=rank * pairs, total_pairs=world * pairs, max_finished=2 * n)
    lazy = mode == "lazy"

    def ev():
        return torch.cuda.Event(enable_timing=True)

    out = {"generations": []}
    for gen in range(generations):
        states = env.reset_all(lazy=True) if lazy else env.reset_all()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0, t1, t2 = ev(), ev(), ev()
        t0.record()
        for _ in range(steps):
            actions = agent.step(states)
            states, rewards, dones, _ = env.step_lazy(actions) if lazy else env.step(actions)
            agent.store_async(rewards, dones)
        t1.record()
        n_fin = int(agent._counters[0].item())
        theta_before = agent.network.theta_packed().clone()
        agent.train()
        t2.record()
        torch.cuda.synchronize()
        stats = par.all_reduce_episode_stats(par.stats_vector(env.stats()))
        theta = agent.network.theta_packed()
        # consistency: every rank must hold identical parameters after the update
        h = torch.stack([theta.double().sum(), (theta.double() * torch.arange(theta.numel(), device=dev)).sum()])
        hmin, hmax = h.clone(), h.clone()
        ms = torch.tensor([t0.elapsed_time(t1), t1.elapsed_time(t2)], device=dev, dtype=torch.float64)
        fin = torch.tensor([n_fin], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(hmin, op=dist.ReduceOp.MIN)
            dist.all_reduce(hmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(fin, op=dist.ReduceOp.SUM)
        out["generations"].append({
            "rollout_ms": ms[0].item(), "ms_per_step": ms[0].item() / steps, "train_ms": ms[1].item(),
            "env_steps_per_sec_with_policy": total * steps / (ms[0].item() * 1e-3),
            "episodes_ranked": fin.item(), "mean_return": stats["mean_return"],
            "theta_step_norm": float((theta - theta_before).norm()), "theta_norm": float(theta.norm()),
            "parameters_identical_on_all_ranks": bool(torch.equal(hmin, hmax)),
        })
    net = agent.network
    out.update(n_gpus=world, total_envs=total, envs_per_gpu=n, eval_envs_per_gpu=n_eval, steps=steps, window=W,
               mode=mode, network_shape=list(net.shape), policy_params=sum(w.numel() + b.numel() for w, b in
                                                                           zip(net.weight_layers, net.bias_layers)),
               perturbation_bytes_per_gpu=net._eps.numel() * 2, workload=workload,
               perturbation_storage="fp16 (the reference draws f32; parity shown for fp16-representable eps)",
               kernels_per_step=["fe_es_forward", "fe_lazy_kernel", "fe_es_store"],
               fitness_gather="finished-episode lists (global env id, return), all_gather over NCCL")
    del env, agent, series
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs-per-gpu", type=int, default=524288)
    ap.add_argument("--eval-envs-per-gpu", type=int, default=0)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--generations", type=int, default=3)
    ap.add_argument("--window", type=int, default=60)
    ap.add_argument("--hidden", type=int, nargs="*", default=[8])
    ap.add_argument("--workload", default="c4", choices=["c2", "c4"])
    ap.add_argument("--mode", default="lazy", choices=["lazy", "dense"])
    ap.add_argument("--sigma", type=float, default=0.05)
    ap.add_argument("--lr", type=float, default=0.01)
    args = ap.parse_args()

    from finenvs_b200 import parallel as par

    rank, world, local_rank = par.init_distributed("nccl")
    torch.cuda.set_device(local_rank)
    out = run(rank, world, local_rank, args.envs_per_gpu, args.eval_envs_per_gpu, args.steps, args.generations, args.window,
              tuple(args.hidden), args.workload, args.mode, args.sigma, args.lr)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
