"""Device-resident step time across observation dtypes / windows / kernels (auto vs forced pipe): tools/f64_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, bench
from finenvs_b200.data import loader
from finenvs_b200.environments import TimeSeriesEnv
cases = [(60, 1 << 20, torch.float64, "auto"), (60, 1 << 20, torch.float64, "pipe"), (60, 1 << 20, torch.float32, "auto"),
         (48, 1310720, torch.float64, "auto"), (100, 314560, torch.float64, "auto"), (100, 314560, torch.float64, "pipe"),
         (104, 604160, torch.float32, "auto"), (104, 604160, torch.float32, "pipe"),
         (128, 491520, torch.float32, "auto"), (128, 491520, torch.float32, "pipe")]
for W, N, dt, variant in cases:
    prices, seg_start, seg_len, _ = bench.make_series("c2", W)
    if True:
        series = loader.stage_series(prices, seg_start, seg_len, W, "cuda:0", dt)
        env = TimeSeriesEnv("p", num_intervals=W, device_id=0, series=series, num_envs=N, seed=3, random_reset="all", random_offset=True,
                            obs_dtype=dt, variant=variant)
        env.reset()
        a = [torch.rand((N, 1), device="cuda") * 2 - 1 for _ in range(4)]
        obs = torch.empty((N, W, 5), dtype=dt, device="cuda"); r = torch.empty(N, dtype=dt, device="cuda"); d = torch.empty(N, dtype=torch.int32, device="cuda")
        for i in range(5): env.step_into(a[i % 4], obs, r, d)
        ms = []
        for b in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for i in range(20): env.step_into(a[i % 4], obs, r, d)
            e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1) / 20)
        gb = obs.numel() * obs.element_size() / 1e9
        print(W, N, dt, env.kernel_name(), "ms", round(sorted(ms)[2], 4), "obs GB", round(gb, 3), "write TB/s", round(gb / sorted(ms)[2], 3))
        del env, series, obs
