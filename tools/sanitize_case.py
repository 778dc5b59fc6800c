"""Smallest end-to-end case for compute-sanitizer (memcheck / racecheck / synccheck) touching every kernel of the
library: all step variants x both dtypes (ragged tail tiles, resets, evaluate mode), portfolio, lazy step +
materialise, the three ES forward kernels + perturb / gradient / store, PPO returns, captured rollout, host step.
Plain run: a smoke test of every launch path.  Under compute-sanitizer (where the pool allows it):
    compute-sanitizer --tool memcheck python tools/sanitize_case.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from parity_utils import gbm_ohlc  # noqa: E402
from finenvs_b200.agents.ES import EvoAgent  # noqa: E402
from finenvs_b200.agents.networks import ParallelMLP  # noqa: E402
from finenvs_b200.agents.PPO import Buffer  # noqa: E402
from finenvs_b200.data import loader  # noqa: E402
from finenvs_b200.environments import TimeSeriesEnv  # noqa: E402

rng = np.random.default_rng(0)
bars, days = 12, 9
for W in (5, 8, 60):
    prices = np.round(gbm_ohlc(rng, W + bars * days, 0.05), 4)
    firsts = W + bars * np.arange(days)
    for dtype in (torch.float32, torch.float64):
        series = loader.stage_series(prices, firsts - W, np.full(days, W + bars, np.int32), W, "cuda:0", dtype)
        for variant in ("tile", "direct", "pipe", "gather", "split"):
            if variant == "gather" and series.obs_table() is None:   # 5 * W * itemsize is not a legal TMA row
                continue
            for kw in (dict(random_reset="all", random_offset=True, track_stats=True), dict(evaluate=True)):
                env = TimeSeriesEnv("san", num_intervals=W, series=series, num_envs=None if "evaluate" in kw else 203,
                                    seed=1, obs_dtype=dtype, variant=variant, **kw)
                env.reset()
                for t in range(20):
                    a = torch.rand((env.num_envs, 1), device="cuda") * 2 - 1
                    env.step(a)
                env.reset_all() if "evaluate" not in kw else None

# portfolio (A = 3), host step (pinned -> zero-copy with the portfolio kernel; pageable -> chunked copies), captured rollout
W, A = 8, 3
pp = np.stack([np.round(gbm_ohlc(rng, W + bars * days, 0.05, s0=30.0 + 9 * a), 4) for a in range(A)], axis=1)
series3 = loader.stage_series(pp, (W + bars * np.arange(days)) - W, np.full(days, W + bars, np.int32), W, "cuda:0", torch.float32)
env = TimeSeriesEnv("port", num_intervals=W, series=series3, num_envs=77, seed=2, random_reset="all", random_offset=True,
                    track_stats=True)
env.reset()
for t in range(20):
    a = torch.rand((77, A)) * 2 - 1
    env.step(a.cuda())
    env.step_host(a.pin_memory() if t % 2 else a)

W = 12
prices = np.round(gbm_ohlc(rng, W + bars * days, 0.05), 4)
series = loader.stage_series(prices, (W + bars * np.arange(days)) - W, np.full(days, W + bars, np.int32), W, "cuda:0", torch.float32)
env = TimeSeriesEnv("cap", num_intervals=W, series=series, num_envs=301, seed=3, random_reset="all", random_offset=True)
roll = env.capture_rollout(lambda obs: torch.tanh(obs[:, -1, 3:4] * 30.0), 5)
roll.replay(); roll.replay()
a = torch.rand((301, 1)) * 2 - 1
env.step_host(a.pin_memory()); env.step_host(a)
# host step of the persistent kernels with pinned buffers (zero-copy reads, claim-ahead bookkeepers), both wire formats of the
# dones: gather packs and sends them itself, pipe through the staging buffer + fe_flush_bits_kernel
for variant in ("gather", "pipe"):
    env = TimeSeriesEnv("host", num_intervals=W, series=series, num_envs=1301, seed=3, random_reset="all", random_offset=True,
                        variant=variant, track_stats=True)
    a = (torch.rand((1301, 1)) * 2 - 1).pin_memory()
    for t in range(6):
        env.step_host(a, packed_dones=bool(t % 2))

# ES: fast (60-8-1), streaming (60-16-4-2), generic (400-6-1 needs W = 80) forward kernels, lazy and dense, eval envs
for shape, Wn in (((60, 8, 1), 12), ((60, 16, 4, 2), 12), ((400, 6, 1), 80)):
    pr = np.round(gbm_ohlc(rng, Wn + 100 * 4, 0.05), 4)
    ser = loader.stage_series(pr, (Wn + 100 * np.arange(4)) - Wn, np.full(4, Wn + 100, np.int32), Wn, "cuda:0", torch.float32)
    N, E = 206, 6
    env = TimeSeriesEnv("es", num_intervals=Wn, series=ser, num_envs=N, seed=4, random_reset="all", random_offset=True,
                        flat_obs=True, num_eval_envs=E)
    torch.manual_seed(0)
    if shape[-1] == 1:
        agent = EvoAgent(env.get_env_args(), hidden_dims=tuple(shape[1:-1]), write_to_csv=False, seed=5)
        states = env.reset_all(lazy=True)
        for t in range(25):
            actions = agent.step(states)
            states, rewards, dones, _ = env.step_lazy(actions)
            agent.store(rewards, dones)
        states.materialize()
        agent.train()
    else:
        net = ParallelMLP(N, E, shape, device_id=0, seed=6)
        net.perturb_parameters()
        lo = env.reset_lazy()
        net.forward(lo); net.forward(lo.materialize())
        net.update_parameters(torch.rand(N, device="cuda"))

# PPO buffer: zero-copy observation hand-off + returns kernel
env = TimeSeriesEnv("ppo", num_intervals=12, series=series, num_envs=130, seed=7)
buf = Buffer(4, 0.99, 0, capacity=8)
buf.bind_env(env)
s0 = env.reset()
for t in range(10):
    act = torch.rand((130, 1), device="cuda") * 2 - 1
    s1, r, d, _ = env.step(act)
    buf.store(s0, act, r.view(-1, 1), d.view(-1, 1), torch.zeros(130, 1, device="cuda"), torch.rand(130, 1, device="cuda"))
    s0 = s1
buf.prepare_training_data(torch.rand(130, 1, device="cuda"))
torch.cuda.synchronize()
print("sanitize_case: ok")
