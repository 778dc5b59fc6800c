"""Smallest end-to-end case for compute-sanitizer (memcheck / racecheck / synccheck): both kernel
variants, both dtypes, ragged tail blocks, resets, evaluate mode.  Run under gpurun:
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
from finenvs_b200.data import loader  # noqa: E402
from finenvs_b200.environments import TimeSeriesEnv  # noqa: E402

rng = np.random.default_rng(0)
for W in (5, 60):
    bars, days = 12, 9
    prices = np.round(gbm_ohlc(rng, W + bars * days, 0.05), 4)
    firsts = W + bars * np.arange(days)
    for dtype in (torch.float32, torch.float64):
        series = loader.stage_series(prices, firsts - W, np.full(days, W + bars, np.int32), W, "cuda:0", dtype)
        for variant in ("tile", "direct"):
            for kw in (dict(random_reset="all", random_offset=True, track_stats=True), dict(evaluate=True)):
                env = TimeSeriesEnv("san", num_intervals=W, series=series, num_envs=None if "evaluate" in kw else 203,
                                    seed=1, obs_dtype=dtype, variant=variant, **kw)
                env.reset()
                for t in range(30):
                    a = torch.rand((env.num_envs, 1), device="cuda") * 2 - 1
                    env.step(a)
                env.reset_all() if "evaluate" not in kw else None
torch.cuda.synchronize()
print("sanitize_case: ok")
