#!/usr/bin/env python
"""bench.py — env-steps/s of the fused trading-env step kernel (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c4]

A "step" is one pass of the hot path over one batch of envs: one fe_step launch over this GPU's
envs.  Default workload = BASELINE config 2 (single asset, 1 Mi envs per GPU, W=60, synthetic GBM
daily bars); `--workload c4` swaps in the 10 M-row minute-bar series (bigger than L2).  N>1 is
launched by torchrun, one rank per GPU; envs are sharded (weak scaling: 1 Mi envs per GPU), the
series is replicated, the step needs no collective.  Rank 0 prints ONE JSON line.

value  = device-resident throughput (actions already in HBM), CUDA events, max over ranks.
e2e    = same metric through the host-buffer C-ABI call fe_step_host (TimeSeriesEnv.step_host):
         per step 4N bytes of actions pinned-host -> HBM and 8N bytes (rewards f32 + dones i32)
         HBM -> pinned host; the observation stays in HBM for the policy.
roofline = algorithmic bytes per launch (2264 B per env-step at A=1, W=60, f32 obs)
         / average launch duration, against MEASURED_PEAKS.json hbm_gbs.
cpu_baseline / --impl reference = the CPU oracle (a C port of the reference's algorithm; the
         reference itself is torch-eager Python and cannot travel to the GPU box) on all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"
SERIES_SEED = 20260101
ACTION_SEED = 1234


# ------------------------------------------------------------------------------ workloads ----
def algorithmic_bytes_per_env_step(W: int, A: int = 1, obs_bytes: int = 4) -> int:
    """window read + obs write + OHLC row + state r/w + tables + action + reward + done."""
    window_read = W * 4 * A * obs_bytes
    obs_write = W * 5 * A * obs_bytes
    ohlc = 32 * A
    cash = 8
    per_asset_state = A * (8 + 8 + 16)   # long, short (f32 r+w), margin (f64 r+w)
    pointer = 8
    tables = 12                           # segment id 4 r, seg_start 8 r (seg_len shares the line)
    return window_read + obs_write + ohlc + cash + per_asset_state + pointer + tables + 4 * A + obs_bytes + 4


def make_series(workload: str, W: int):
    """Synthetic GBM OHLC (SURVEY.md §8d recipe) + segment table."""
    from finenvs_b200.data import loader
    from parity_utils import gbm_ohlc

    rng = np.random.default_rng(SERIES_SEED)
    if workload == "c4":     # minute bars, 10 M rows, 390-bar segments
        T, bars, sigma = 10_000_000, 390, 0.0005
    else:                    # c2 / c1 / c3: daily bars, 1024 segments x 252 bars
        T, bars, sigma = 1024 * 252, 252, 0.01
    if workload == "c3":     # 30 independent GBM assets, time-major (T, A, 4)
        prices = np.stack([np.round(gbm_ohlc(rng, T, sigma, s0=20.0 + 7 * a), 4) for a in range(WORKLOAD_ASSETS["c3"])], axis=1)
    else:
        prices = np.round(gbm_ohlc(rng, T, sigma), 4)
    seg_start, seg_len = loader.regular_segments(T, bars, W)
    return prices, seg_start, seg_len, {"rows": T, "bars_per_segment": bars, "sigma": sigma}


WORKLOAD_ASSETS = {"c2": 1, "c4": 1, "c3": 30}
WORKLOAD_DEFAULTS = {"c2": (1 << 20, 60), "c4": (1 << 20, 60), "c3": (65536, 128)}   # (envs per GPU, window)

WORKLOAD_NAMES = {
    "c3": "30-asset portfolio-allocation env, 65536 envs, 128-step window, transaction costs, 1 B200",
    "c2": "single-asset env, 1M envs, 60-step window, fused step kernel on 1 B200 vs reference",
    "c4": "synthetic minute-bar series of 10M timesteps, 1M envs per GPU with random start offsets, env-sharded",
}


# ------------------------------------------------------------------------------ clocks -------
class ClockSampler:
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz, self._stop = [], set(), None, threading.Event()
        self._t = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None

    def _run(self):
        nv = self._nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self._nv is not None:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._t is not None:
            self._t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------ CPU legs -----
def cpu_oracle_throughput(W: int, workload: str, sample_envs: int, steps: int, warmup: int, budget_s: float | None):
    """The oracle port stepped on all host threads over a bounded sample of the workload.
    Returns (env-steps/s, seconds per step, threads, steps timed)."""
    from oracle import oracle as orc

    prices, seg_start, seg_len, _ = make_series(workload, W)
    fs = orc.series_from_prices(prices, seg_start, seg_len, W)
    env = orc.OracleEnv(fs, num_envs=sample_envs, seed=ACTION_SEED, reset_mode=orc.RESET_ALL, random_offset=True,
                        out_f64=False)
    threads = orc.lib().feo_num_threads()
    rng = np.random.default_rng(ACTION_SEED)
    A = WORKLOAD_ASSETS[workload]
    acts = [rng.uniform(-1, 1, sample_envs * A).astype(np.float32) for _ in range(4)]
    obs = np.empty((sample_envs, W, 5 * A), np.float32)
    rewards = np.empty(sample_envs, np.float32)
    dones = np.empty(sample_envs, np.int32)
    import ctypes as C

    def one(i):
        env.step_count += 1
        fn = orc.lib().feo_step_multi if env.multi else orc.lib().feo_step
        fn(C.byref(env.p), C.byref(env.s), C.byref(env.st), orc._p(acts[i % 4]), orc._p(obs),
           orc._p(rewards), orc._p(dones), env.step_count, None)

    for i in range(warmup):
        one(i)
    t0 = time.perf_counter()
    done_steps = 0
    for i in range(steps):
        one(i)
        done_steps += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s and done_steps >= 3:
            break
    dt = time.perf_counter() - t0
    return sample_envs * done_steps / dt, dt / done_steps, threads, done_steps


def run_reference_arm(args, rank: int):
    """--impl reference: the reference's algorithm on the box's host cores (oracle port, all threads)."""
    if rank != 0:
        return
    sample = min(args.envs, 262144 // WORKLOAD_ASSETS[args.workload] // (2 if args.workload == "c3" else 1))
    v, sps, threads, n = cpu_oracle_throughput(args.window, args.workload, sample, args.steps, args.warmup, None)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": n,
        "warmup": args.warmup, "ms_per_step": sps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32/f64", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} envs per step, {n} steps, oracle/fe_oracle.c (OpenMP, {threads} threads)"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world: int):
    return {"workload": WORKLOAD_NAMES[args.workload], "envs_per_gpu": args.envs, "total_envs": args.envs * world,
            "window": args.window, "assets": WORKLOAD_ASSETS[args.workload], "obs_dtype": "float32", "reset": "all envs redraw (segment, offset), Philox",
            "parallelism": f"env-sharded x{world}, series replicated, no per-step collective",
            "l2": "per-step working set (obs >= 1.2 GB + state) >> 126 MB L2: no flush needed"}


# ------------------------------------------------------------------------------ GPU arm ------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4"])
    ap.add_argument("--envs", type=int, default=None, help="envs per GPU (default: 1 Mi; c3: 65536)")
    ap.add_argument("--window", type=int, default=None, help="default 60; c3: 128")
    ap.add_argument("--variant", default="auto", choices=["auto", "tile", "direct", "pipe", "scatter", "split", "rows"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin the rank to its GPU's NUMA-local CPUs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    args.envs = args.envs or WORKLOAD_DEFAULTS[args.workload][0]
    args.window = args.window or WORKLOAD_DEFAULTS[args.workload][1]
    A = WORKLOAD_ASSETS[args.workload]

    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist

    from finenvs_b200 import parallel as par
    from finenvs_b200.data import loader

    prev_affinity = None
    if not args.no_numa_bind:   # before the CUDA context and any pinned allocation exist
        prev_affinity = par.bind_to_gpu_numa(int(os.environ.get("LOCAL_RANK", "0")))
    rank, world, local_rank = par.init_distributed("nccl")
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    W, N = args.window, args.envs
    total = N * world

    prices, seg_start, seg_len, meta = make_series(args.workload, W)
    series = loader.stage_series(prices, seg_start, seg_len, W, dev, torch.float32)
    env = par.make_sharded_env(total, rank, world, "bench", num_intervals=W, device_id=local_rank, series=series,
                               seed=ACTION_SEED, random_reset="all", random_offset=True, variant=args.variant)
    assert env.num_envs == N

    # inputs resident in HBM: a ring of pre-generated action batches
    g = torch.Generator(device=dev).manual_seed(ACTION_SEED + rank)
    ring = [torch.rand((N, A), generator=g, device=dev) * 2 - 1 for _ in range(8)]
    obs = torch.empty((N, W, 5 * A), dtype=torch.float32, device=dev)
    rewards = torch.empty(N, dtype=torch.float32, device=dev)
    dones = torch.empty(N, dtype=torch.int32, device=dev)
    ring_host = [a.cpu().pin_memory() for a in ring[:4]]

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    stream = torch.cuda.current_stream()
    sampler = ClockSampler(local_rank)
    with sampler:
        # ---- device-resident: `value` and the kernel's roofline --------------------------------
        for i in range(args.warmup):
            env.step_into(ring[i % 8], obs, rewards, dones)
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(args.steps):
            env.step_into(ring[i % 8], obs, rewards, dones)
        e1.record(stream)
        sync_all()
        dev_ms = max_over_ranks(e0.elapsed_time(e1))
        n_done = int(dones.sum().item())

        # ---- end to end through the host-buffer C-ABI call ----------------------------------------
        for i in range(args.warmup):
            env.step_host(ring_host[i % 4])
        sync_all()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_wall = time.perf_counter()
        e2.record(stream)
        checksum = 0.0
        for i in range(args.steps):
            _, r_h, d_h, _ = env.step_host(ring_host[i % 4])
            checksum += float(r_h[0]) + int(d_h[0])   # the host reads the step's result
        e3.record(stream)
        sync_all()
        e2e_ms = max_over_ranks(max(e2.elapsed_time(e3), (time.perf_counter() - t_wall) * 1e3))

    value = total * args.steps / (dev_ms * 1e-3)
    e2e_value = total * args.steps / (e2e_ms * 1e-3)
    bytes_per_launch = algorithmic_bytes_per_env_step(W, A) * N
    kernel_s = dev_ms * 1e-3 / args.steps   # only fe_step launches sit between the two events
    achieved = bytes_per_launch / kernel_s / 1e9
    peak, peak_src = 6650.0, "fallback"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak, peak_src = float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        pass
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get(f"{args.workload}_w{W}_n{N}")
    except Exception:
        pass

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32/f64", "data": "synthetic", "config": {**workload_config(args, world), **meta},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": f"of {peak_src}",
                     "algorithmic_bytes_per_env_step": algorithmic_bytes_per_env_step(W, A),
                     "kernel": env.kernel_name(),
                     "kernel_ms": kernel_s * 1e3,
                     # what DRAM actually carried (ncu) over the same launch time: when the series is L2-resident the
                     # window reads never reach DRAM, so `frac` (algorithmic bytes) can exceed 1 while this stays below
                     "dram_gbs": (traffic / kernel_s / 1e9) if traffic else None,
                     "dram_frac": (traffic / kernel_s / 1e9 / peak) if traffic else None},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 4 * N * A, "d2h_bytes_per_step": 8 * N,
                "ms_per_step": e2e_ms / args.steps, "api": "TimeSeriesEnv.step_host -> fe_step_host (pinned host buffers)"},
        # kernels of this library inside the two timed regions (device + e2e legs): one per step, two where the step is
        # a bookkeeping + a streaming launch (portfolio, split); with pinned buffers the host step is the same launch(es)
        "gpu_launches": 2 * args.steps * (2 if " + " in env.kernel_name() else 1),
        "clocks": sampler.summary(),
        "episodes_finished_last_step": n_done,
    }
    line["config"]["numa_bound_cpus"] = len(os.sched_getaffinity(0)) if prev_affinity is not None else None
    if prev_affinity is not None:
        os.sched_setaffinity(0, prev_affinity)   # the CPU baseline uses every host core
    if world == 1 and not args.no_cpu_baseline:
        sample = min(N, 262144 // A // (2 if A > 1 else 1))
        v, sps, threads, n = cpu_oracle_throughput(W, args.workload, sample, 10_000, 3, 12.0)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"{sample} envs per step, {n} steps (~12 s), oracle/fe_oracle.c, OpenMP {threads} threads"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
