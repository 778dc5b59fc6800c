#!/usr/bin/env python
"""Phase clocks of the gather kernel (experiment build with -DFE_GATHER_CLOCKS, tools/build_gather_variants.sh ...c):
FINENVS_B200_LIB=finenvs_b200/libfe_ga_b6m12c.so python tools/gather_clocks.py"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import bench  # noqa: E402
from finenvs_b200 import _lib  # noqa: E402
from finenvs_b200.data import loader  # noqa: E402
from finenvs_b200.environments import TimeSeriesEnv  # noqa: E402

W, N = 60, 1 << 20
prices, seg_start, seg_len, _ = bench.make_series("c2", W)
series = loader.stage_series(prices, seg_start, seg_len, W, "cuda:0", torch.float32)
env = TimeSeriesEnv("clk", num_intervals=W, device_id=0, series=series, num_envs=N, seed=1, random_reset="all", random_offset=True)
print(env.kernel_name())
acts = [torch.rand((N, 1), device="cuda") * 2 - 1 for _ in range(4)]
obs = torch.empty((N, W, 5), device="cuda"); rew = torch.empty(N, device="cuda"); dn = torch.empty(N, dtype=torch.int32, device="cuda")
for i in range(10):
    env.step_into(acts[i % 4], obs, rew, dn)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(50):
    env.step_into(acts[i % 4], obs, rew, dn)
e1.record()
torch.cuda.synchronize()
print(f"{e0.elapsed_time(e1) / 50:.4f} ms per step")
e0.record()
for i in range(50):
    env._observe_into(obs)
e1.record()
torch.cuda.synchronize()
print(f"{e0.elapsed_time(e1) / 50:.4f} ms per reset() [fe_observe: bookkeeping reduced to reading seg / ptr / shares / close]")
for i in range(3):
    env.step_into(acts[i % 4], obs, rew, dn)
torch.cuda.synchronize()
out = (C.c_ulonglong * 16)()
L = _lib.lib()
L.fe_debug_gather_clocks.argtypes = [C.POINTER(C.c_ulonglong)]
assert L.fe_debug_gather_clocks(out) == 0
v = [int(x) for x in out]
n, nb = max(v[6], 1), max(v[10], 1)
print(f"mover 0 of block 0, {n} units: per unit  wait_full {v[0] / n:.0f}  pf {v[1] / n:.0f}  fence+sync {v[11] / n:.0f}  store+release {v[2] / n:.0f}  "
      f"wait_store_read {v[3] / n:.0f}  wait_desc {v[4] / n:.0f}  tma_issue {v[5] / n:.0f}   total {(sum(v[:6]) + v[11] + v[12]) / n:.0f} cycles")
print(f"mover 0 elapsed {v[13]} cycles = {v[13] / n:.0f} per unit")
print(f"bookkeeper 0 of block 0, {nb} tiles: per tile  wait_free {v[8] / nb:.0f}  env_step {v[9] / nb:.0f} cycles")

blk = (C.c_ulonglong * 480)()
L.fe_debug_gather_blocks.argtypes = [C.POINTER(C.c_ulonglong)]
assert L.fe_debug_gather_blocks(blk) == 0
import numpy as np
b = np.array([int(x) for x in blk], dtype=np.int64).reshape(160, 3)[:148]
t0 = b[:, 0].min()
start, end, smid = (b[:, 0] - t0) / 1e3, (b[:, 1] - t0) / 1e3, b[:, 2]
dur = end - start
print(f"blocks: start spread {start.max():.1f} us; duration min {dur.min():.1f}  p10 {np.percentile(dur, 10):.1f}  median {np.median(dur):.1f}  "
      f"p90 {np.percentile(dur, 90):.1f}  max {dur.max():.1f} us; last end {end.max():.1f} us; block 0: {dur[0]:.1f} us on SM {smid[0]}")
order = np.argsort(dur)
print("fastest SMs:", [(int(smid[i]), round(float(dur[i]), 1)) for i in order[:6]], " slowest:", [(int(smid[i]), round(float(dur[i]), 1)) for i in order[-6:]])
