"""Launch sequence for `ncu --metrics gpu__time_duration.sum -k regex:fe_gather`: 6 device-resident steps, 6 host-buffer steps with
int32 dones, 6 with bit-packed dones (c2, 1 Mi envs) — the kernel's own duration in each mode, free of launch and sync latency."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import bench  # noqa: E402
from finenvs_b200.data import loader  # noqa: E402
from finenvs_b200.environments import TimeSeriesEnv  # noqa: E402

N, W = 1 << 20, 60
prices, seg_start, seg_len, _ = bench.make_series("c2", W)
series = loader.stage_series(prices, seg_start, seg_len, W, "cuda:0", torch.float32)
env = TimeSeriesEnv("probe", num_intervals=W, device_id=0, series=series, num_envs=N, seed=3, random_reset="all", random_offset=True)
env.reset()
a = torch.rand((N, 1), device="cuda:0") * 2 - 1
ah = a.cpu().pin_memory()
obs = torch.empty((N, W, 5), dtype=torch.float32, device="cuda:0")
r = torch.empty(N, dtype=torch.float32, device="cuda:0")
d = torch.empty(N, dtype=torch.int32, device="cuda:0")
for _ in range(6):
    env.step_into(a, obs, r, d)
torch.cuda.synchronize()
for _ in range(6):
    env.step_host(ah)
for _ in range(6):
    env.step_host(ah, packed_dones=True)
print("done")
