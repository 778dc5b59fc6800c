"""TimeSeriesEnv — drop-in for the reference's finenvs/environments/time_series_env.py:14-536.

Same constructor arguments, attributes and `reset()` / `step(actions)` / `get_env_args()` contract
(torch tensors in and out, so finenvs/agents run unchanged), but the series is staged once into
HBM as a flat table and every `step` is ONE launch of the sm_100a kernel in csrc/fe_step.cu,
reached through the C ABI of include/finenvs_b200.h.  There is no CPU path.

Keyword-only arguments after `device_id` are extensions; their defaults
reproduce the reference's behaviour, except `obs_dtype` (float32; pass torch.float64 for the
reference's dtype) and the redraw RNG (counter-based Philox instead of torch's global generator).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from .. import _lib
from ..base_object import BaseObject
from ..data import loader
from ..device_utils import require_cuda_device

try:  # gym is metadata only (:218-234); the reference needs it, we do not
    from gym import spaces as _spaces
except Exception:  # pragma: no cover - gym is not installed in the target image
    class _Box:
        def __init__(self, low, high, shape=None, dtype=None):
            self.low, self.high, self.dtype = low, high, dtype
            self.shape = shape if shape is not None else np.shape(low)

    class _spaces:  # type: ignore
        Box = _Box

_RESET_MODES = {"keep": _lib.RESET_KEEP, "last": _lib.RESET_LAST, "all": _lib.RESET_ALL}
_VARIANTS = {"auto": _lib.VARIANT_AUTO, "tile": _lib.VARIANT_TILE, "direct": _lib.VARIANT_DIRECT,
             "portfolio": _lib.VARIANT_PORTFOLIO, "pipe": _lib.VARIANT_PIPE, "split": _lib.VARIANT_SPLIT,
             "gather": _lib.VARIANT_GATHER}


class LazyObs:
    """An observation that has not been materialised: obs[i, j, 0:4] = logret[row0[i] + j], obs[i, j, 4] = posfeat[i]
    (time_series_env.py:428-445).  12 bytes per env instead of W x 20; consumers that read the window straight from
    the staged series (finenvs_b200.agents.networks.ParallelMLP.forward) never need the tensor."""

    __slots__ = ("env", "row0", "posfeat")

    def __init__(self, env: "TimeSeriesEnv", row0: torch.Tensor, posfeat: torch.Tensor):
        self.env, self.row0, self.posfeat = env, row0, posfeat

    @property
    def logret(self) -> torch.Tensor:
        return self.env.series.logret

    @property
    def window(self) -> int:
        return self.env.num_intervals

    @property
    def shape(self):
        e = self.env
        return (e.num_envs, e.num_intervals * e.num_obs) if e.flat_obs else (e.num_envs, e.num_intervals, e.num_obs)

    def materialize(self) -> torch.Tensor:
        """The tensor step() would have returned for this observation."""
        e = self.env
        obs = torch.empty(self.shape, dtype=e.obs_dtype, device=e._dev)
        _lib.check(e._L.fe_materialize(e._pp, e._ps, self.row0.data_ptr(), self.posfeat.data_ptr(), obs.data_ptr(), e._stream()),
                   "fe_materialize")
        return obs


class CapturedRollout:
    """`num_steps` iterations of `actions = policy(obs); obs, rewards, dones = env.step(actions)` recorded ONCE in a
    CUDA graph and replayed with a single launch.  The loop it replaces is the reference's evaluation
    loop (examples/time_series/PPO_LSTM_testing_SPY.py:43-52), whose cost for a few thousand envs is Python and launch
    overhead, not the step.  The step ordinal lives in device memory (fe_step_captured), so a replay continues the
    env exactly where eager stepping would be: results are identical to calling env.step() num_steps times.

    Static tensors (overwritten by every replay): `obs` (N, W, 5) current observation, `rewards` (num_steps, N),
    `dones` (num_steps, N) int32, `actions` (num_steps, N, num_acts)."""

    def __init__(self, env: "TimeSeriesEnv", policy, num_steps: int):
        if num_steps < 1:
            raise ValueError("num_steps must be >= 1")
        self.env, self.policy, self.num_steps = env, policy, int(num_steps)
        dev, N = env._dev, env.num_envs
        shape = (N, env.num_intervals * env.num_obs) if env.flat_obs else (N, env.num_intervals, env.num_obs)
        self.obs = torch.empty(shape, dtype=env.obs_dtype, device=dev)
        self.rewards = torch.empty((num_steps, N), dtype=env.obs_dtype, device=dev)
        self.dones = torch.empty((num_steps, N), dtype=torch.int32, device=dev)
        self.actions = torch.empty((num_steps, N, env.num_acts), dtype=torch.float32, device=dev)
        self._counter = torch.zeros(1, dtype=torch.int64, device=dev)
        self.graph = torch.cuda.CUDAGraph()
        # warm-up on a side stream (lazy initialisation inside the policy / the library must not be captured), on a
        # snapshot of the env state that is restored afterwards
        snap = env.state_snapshot()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            env._observe_into(self.obs)
            self._counter.fill_(env.step_count)
            self._record(1)
        torch.cuda.current_stream(dev).wait_stream(side)
        env.load_state_snapshot(snap)
        with torch.cuda.graph(self.graph):
            self._record(self.num_steps)
        env.load_state_snapshot(snap)   # capture does not execute, but keep the contract explicit

    def _record(self, steps: int) -> None:
        env = self.env
        for t in range(steps):
            a = self.policy(self.obs)
            self.actions[t].copy_(a.reshape(env.num_envs, env.num_acts))
            _lib.check(
                env._L.fe_step_captured(env._pp, env._ps, env._pst, self.actions[t].data_ptr(), self.obs.data_ptr(),
                                        self.rewards[t].data_ptr(), self.dones[t].data_ptr(),
                                        env._stats.data_ptr() if env._stats is not None else None,
                                        self._counter.data_ptr(), env._stream()),
                "fe_step_captured",
            )

    def replay(self, refresh_obs: Optional[bool] = None):
        """Run the captured num_steps steps.  `refresh_obs=True` first rebuilds `obs` from the env's current state
        (what env.reset() returns); False continues from the observation the previous replay left (what the eager
        loop does: the observation returned by a step is built before the auto-reset, :321).  Default: refresh only
        if this is the first replay or something else stepped / reset the env in between."""
        env = self.env
        if refresh_obs is None:
            refresh_obs = getattr(self, "_resume_at", None) != (env.step_count, env._epoch)
        if refresh_obs:
            env._observe_into(self.obs)
        self._counter.fill_(env.step_count)
        self.graph.replay()
        env.step_count += self.num_steps
        self._resume_at = (env.step_count, env._epoch)
        return self.obs, self.rewards, self.dones


class TimeSeriesEnv(BaseObject):
    def __init__(
        self,
        instrument_name: str,
        dataset_key: str = "dummy",
        num_intervals: int = 390,
        max_shares: int = 5,
        starting_balance: float = 10000,
        per_share_commission: float = 0.01,
        initial_margin_requirement: float = 1.5,
        maintenance_margin_requirement: float = 0.25,
        evaluate: bool = False,
        device_id: int = 0,
        *,
        num_envs: Optional[int] = None,
        obs_dtype: torch.dtype = torch.float32,
        seed: Optional[int] = None,
        random_reset: Optional[str] = None,
        random_offset: bool = False,
        series: Optional[loader.StagedSeries] = None,
        env_id_base: int = 0,
        total_envs: Optional[int] = None,
        track_stats: bool = False,
        variant: str = "auto",
        flat_obs: bool = False,
        num_eval_envs: Optional[int] = None,
    ):
        """Reference arguments: :15-29.  Extensions:

        num_envs      envs on this device (default: one per trading day, +1 evaluation env when
                      training, :246-257).  Env i starts on segment (global id) mod D.
        obs_dtype     torch.float32 (default) or torch.float64 (reference dtype) for obs and rewards.
        seed          Philox key of the segment redraws; default torch.initial_seed().
        random_reset  "last" (reference training: only the last env redraws its day, :504-513),
                      "keep" (reference evaluate: same day again), "all" (every finished env redraws).
        random_offset redraws also draw the start offset inside the segment.
        series        an already staged series (share one copy between envs / skip the CSV).
        env_id_base, total_envs   this env object is a shard [base, base+num_envs) of a global
                      population (multi-GPU); draws are keyed by global id, so results do not
                      depend on the sharding.
        track_stats   accumulate episode count / return / length on the device (stats()).
        variant       "auto" | "gather" | "pipe" | "tile" | "direct" | "split" | "portfolio" kernel variant (auto: for
                      populations of >= 4 tiles per SM (18 944 envs on a B200) with windows of >= 24 rows the persistent
                      kernels — gather while the series is short enough for its observation-layout table to stay in L2
                      and 5*W*itemsize is a multiple of 16, else pipe; else tile; direct when the window does not fit
                      in shared memory; portfolio whenever the series has more than one asset).
        flat_obs      return observations as (N, W*num_obs) — the 2-D input the ES agent's ParallelMLP needs
                      (parallel_mlp.py:98-103); same memory, only the shape differs.
        num_eval_envs reported in get_env_args() for the ES agent (evo_agent.py:53); the last
                      num_eval_envs envs are the evaluation envs by the reference's convention.
        """
        if obs_dtype not in (torch.float32, torch.float64):
            raise ValueError("obs_dtype must be torch.float32 or torch.float64")
        self.instrument_name = instrument_name
        self.num_intervals = int(num_intervals)
        self.max_shares = max_shares
        self.starting_balance = starting_balance
        self.per_share_commission = per_share_commission
        self.initial_margin_requirement = initial_margin_requirement
        self.maintenance_margin_requirement = maintenance_margin_requirement
        self.log_return_scale_factor = 100
        self.evaluate = evaluate
        self.obs_dtype = obs_dtype
        self.device = require_cuda_device(device_id)
        self._dev = torch.device(self.device)
        self._L = _lib.lib()
        if series is None:
            self.data_dir_name = loader.get_data_dir_name(instrument_name)
            self.file_key = loader.determine_file_key(dataset_key)
            self.filename = loader.find_file_by_key(self.data_dir_name, self.file_key)
            host = loader.read_market_csv(self.filename, self.num_intervals)
            series = loader.stage_series(host.prices, host.seg_start, host.seg_len_raw, self.num_intervals,
                                         self.device, obs_dtype)
        else:
            if series.window != self.num_intervals:
                raise ValueError(f"series was staged for W={series.window}, env asks W={self.num_intervals}")
            if series.device != self._dev:
                raise ValueError(f"series lives on {series.device}, env on {self._dev}")
            if series.logret.dtype != obs_dtype:
                raise ValueError(f"series log-returns are {series.logret.dtype}, obs_dtype is {obs_dtype}")
        self.series = series
        self.num_assets = series.num_assets   # extension: A > 1 = portfolio env sharing one cash account
        self.values_per_interval = 4
        self.set_spaces()
        if random_reset is None:
            random_reset = "keep" if evaluate else "last"
        if random_reset not in _RESET_MODES:
            raise ValueError(f"random_reset must be one of {sorted(_RESET_MODES)}")
        self.random_reset = random_reset
        self.random_offset = bool(random_offset)
        self.seed = int(torch.initial_seed() if seed is None else seed) & 0xFFFFFFFFFFFFFFFF
        self.track_stats = bool(track_stats)
        self.flat_obs = bool(flat_obs)
        self.num_eval_envs = num_eval_envs
        self.set_environment_params(num_envs, env_id_base, total_envs, _VARIANTS[variant])

    # ------------------------------------------------------------------ metadata (:218-243) ----
    def set_spaces(self) -> None:
        self.num_obs = (self.values_per_interval + 1) * self.num_assets
        self.num_acts = self.num_assets
        self.action_space = _spaces.Box(np.ones(self.num_acts) * -1.0, np.ones(self.num_acts) * +1.0, dtype=np.float64)
        self.observation_space = _spaces.Box(
            np.ones((self.num_intervals, self.num_obs)) * -np.inf,
            np.ones((self.num_intervals, self.num_obs)) * +np.inf,
            dtype=np.float64,
        )

    def get_env_args(self) -> Dict:
        env_args = {
            "env_name": self.instrument_name,
            "num_envs": self.num_envs,
            "num_observations": self.num_obs * (self.num_intervals if self.flat_obs else 1),
            "num_actions": self.num_acts,
            "sequence_length": self.num_intervals,
        }
        if self.num_eval_envs is not None:
            env_args["num_eval_envs"] = self.num_eval_envs
        return env_args

    # ------------------------------------------------------------------ state (:245-275) -------
    def set_environment_params(self, num_envs=None, env_id_base=0, total_envs=None, variant=0) -> None:
        D = self.series.num_segments
        if num_envs is None:
            num_envs = D if self.evaluate else D + 1  # :246-257
        N = int(num_envs)
        if N <= 0:
            raise ValueError("num_envs must be positive")
        total = int(N if total_envs is None else total_envs)
        base = int(env_id_base)
        if base < 0 or base + N > total:
            raise ValueError("shard [env_id_base, env_id_base+num_envs) must lie inside total_envs")
        self.num_envs, self.env_id_base, self.total_envs = N, base, total
        dev = self._dev
        gid = torch.arange(base, base + N, device=dev, dtype=torch.int64)
        self._seg = (gid % D).to(torch.int32)
        self._ptr = torch.zeros(N, dtype=torch.int32, device=dev)
        self._cash = torch.full((N,), float(self.starting_balance), dtype=torch.float32, device=dev)
        A = self.num_assets
        self._long = torch.zeros(N * A, dtype=torch.float32, device=dev)
        self._short = torch.zeros(N * A, dtype=torch.float32, device=dev)
        self._margin = torch.zeros(N * A, dtype=torch.float64, device=dev)
        need_ep = self.evaluate or self.track_stats
        self._terminated = torch.zeros(N, dtype=torch.uint8, device=dev) if self.evaluate else None
        self._ep_return = torch.zeros(N, dtype=torch.float32, device=dev) if need_ep else None
        self._ep_len = torch.zeros(N, dtype=torch.int32, device=dev) if self.track_stats else None
        self._stats = torch.zeros(_lib.STATS_BYTES // 8, dtype=torch.int64, device=dev) if need_ep else None
        self._stats_ptr = self._stats.data_ptr() if self._stats is not None else None   # never reallocated (zeroed in place)
        self.step_count = 0
        self._epoch = 0   # bumped by everything that changes the state other than a step (reset_all, snapshots)
        self._params = _lib.FeParams(
            N, base, total, self.series.num_rows, self.num_intervals, D, A, int(self.max_shares),
            float(self.starting_balance), float(self.per_share_commission), float(self.initial_margin_requirement),
            float(self.maintenance_margin_requirement), self.seed, _RESET_MODES[self.random_reset],
            int(self.random_offset), int(self.evaluate), int(self.obs_dtype == torch.float64), variant,
            dev.index if dev.index is not None else torch.cuda.current_device(),
        )
        self._sched = torch.zeros(4, dtype=torch.int32, device=dev)   # gather kernel: tile / arrival counters (FeState.sched)
        s = self.series
        # the observation-layout table is only built for envs that can use it (large populations or variant="gather")
        want_table = variant == _lib.VARIANT_GATHER or (variant == _lib.VARIANT_AUTO and A == 1 and N >= 4096 and self.num_intervals >= 24)
        table = s.obs_table() if want_table else None
        self._cseries = _lib.FeSeries(s.prices.data_ptr(), s.logret.data_ptr(), s.seg_start.data_ptr(), s.seg_len.data_ptr(),
                                      table.data_ptr() if table is not None else None)
        self._cstate = _lib.FeState(
            self._seg.data_ptr(), self._ptr.data_ptr(), self._cash.data_ptr(), self._long.data_ptr(),
            self._short.data_ptr(), self._margin.data_ptr(),
            self._terminated.data_ptr() if self._terminated is not None else None,
            self._ep_return.data_ptr() if self._ep_return is not None else None,
            self._ep_len.data_ptr() if self._ep_len is not None else None,
            self._sched.data_ptr() if table is not None else None,
        )
        self._pp, self._ps, self._pst = C.byref(self._params), C.byref(self._cseries), C.byref(self._cstate)
        if self.random_reset == "last" and base + N == total:
            # :253-257 the extra (evaluation) env starts on a drawn day
            r = _lib.philox(self.seed, total - 1, 0, 1)
            self._seg[-1] = (r[0] * D) >> 32
        elif self.random_reset == "all":
            self._launch_reset_all(redraw=True)

    def reset_evaluation_metrics(self) -> None:
        """:271-275"""
        # in place: CapturedRollout graphs and FeState hold these device addresses
        self._terminated.zero_()
        self._ep_return.zero_()
        self._stats.zero_()

    # reference-named views of the state (same memory; reference shapes (N,1) / (N,))
    @property
    def env_indices(self) -> torch.Tensor:
        return self._seg.long()

    @property
    def env_pointers(self) -> torch.Tensor:
        return self._ptr.long()

    @property
    def env_spots(self) -> torch.Tensor:
        """(N, W) row indices of the window; the reference stores this tensor (:261), here it is derived."""
        return self._ptr.long().unsqueeze(1) + torch.arange(self.num_intervals, device=self._dev)

    @property
    def cash(self) -> torch.Tensor:
        return self._cash.view(-1, 1)

    @property
    def long_shares(self) -> torch.Tensor:
        return self._long.view(self.num_envs, self.num_assets)

    @property
    def short_shares(self) -> torch.Tensor:
        return self._short.view(self.num_envs, self.num_assets)

    @property
    def margin(self) -> torch.Tensor:
        return self._margin.view(self.num_envs, self.num_assets)

    @property
    def terminated_episodes(self) -> torch.Tensor:
        return self._terminated.bool()

    @property
    def episode_returns(self) -> torch.Tensor:
        return self._ep_return

    # ------------------------------------------------------------------ the hot path ------------
    def _stream(self) -> int:
        return torch.cuda.current_stream(self._dev).cuda_stream

    def _new_obs(self, for_reset: bool = False) -> torch.Tensor:
        # a fresh tensor every call: the PPO buffer keeps references to past observations (buffer.py:44-56)
        shape = ((self.num_envs, self.num_intervals * self.num_obs) if self.flat_obs
                 else (self.num_envs, self.num_intervals, self.num_obs))
        ring = getattr(self, "_obs_ring", None)
        if ring is not None:  # a rollout buffer bound with Buffer.bind_env: write straight into its storage
            return ring.next_obs_slot(shape, self.obs_dtype, for_reset)
        return torch.empty(shape, dtype=self.obs_dtype, device=self._dev)

    def reset(self) -> torch.Tensor:
        """:423-435 — materialises the current observation; touches no state (reference semantics)."""
        obs = self._new_obs(for_reset=True)
        _lib.check(self._L.fe_observe(self._pp, self._ps, self._pst, obs.data_ptr(), self._stream()), "fe_observe")
        return obs

    def _prepare_actions(self, actions: torch.Tensor) -> torch.Tensor:
        if not torch.is_tensor(actions):
            raise TypeError("actions must be a torch.Tensor")
        if actions.numel() != self.num_envs * self.num_acts:
            raise ValueError(f"actions must be (num_envs, num_acts) = ({self.num_envs}, {self.num_acts}); "
                             f"got {tuple(actions.shape)}")
        if actions.device != self._dev or actions.dtype != torch.float32 or not actions.is_contiguous():
            actions = actions.to(device=self._dev, dtype=torch.float32, non_blocking=True).contiguous()
        return actions

    def step(self, actions: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, dict]:
        """:277-296 — one kernel launch, no host synchronisation in training mode."""
        actions = self._prepare_actions(actions)
        obs = self._new_obs()
        rewards = torch.empty(self.num_envs, dtype=self.obs_dtype, device=self._dev)
        dones = torch.empty(self.num_envs, dtype=torch.int32, device=self._dev)
        self.step_into(actions, obs, rewards, dones)
        info_dict = self.record_evaluation_metrics() if self.evaluate else {}
        return (obs, rewards, dones, info_dict)

    def step_into(self, actions: torch.Tensor, obs: torch.Tensor, rewards: torch.Tensor, dones: torch.Tensor) -> None:
        """step() into caller-owned output tensors (zero allocation; CUDA-graph capturable)."""
        self.step_count += 1
        _lib.check(
            self._L.fe_step(self._pp, self._ps, self._pst, actions.data_ptr(), obs.data_ptr(), rewards.data_ptr(),
                            dones.data_ptr(), self._stats.data_ptr() if self._stats is not None else None,
                            self.step_count, self._stream()),
            "fe_step",
        )

    def _new_lazy(self) -> LazyObs:
        if self.num_assets != 1:
            raise NotImplementedError("lazy observations are implemented for single-asset envs")
        return LazyObs(self, torch.empty(self.num_envs, dtype=torch.int64, device=self._dev),
                       torch.empty(self.num_envs, dtype=self.obs_dtype, device=self._dev))

    def reset_lazy(self) -> LazyObs:
        """reset() returning a LazyObs handle instead of the tensor."""
        lo = self._new_lazy()
        _lib.check(self._L.fe_observe_lazy(self._pp, self._ps, self._pst, lo.row0.data_ptr(), lo.posfeat.data_ptr(),
                                           self._stream()), "fe_observe_lazy")
        return lo

    def step_lazy(self, actions: torch.Tensor) -> Tuple[LazyObs, torch.Tensor, torch.Tensor, dict]:
        """step() without materialising the observation: same state transition, rewards and dones; the observation
        comes back as a LazyObs handle (`.materialize()` gives exactly the tensor step() returns)."""
        actions = self._prepare_actions(actions)
        lo = self._new_lazy()
        rewards = torch.empty(self.num_envs, dtype=self.obs_dtype, device=self._dev)
        dones = torch.empty(self.num_envs, dtype=torch.int32, device=self._dev)
        self.step_count += 1
        _lib.check(
            self._L.fe_step_lazy(self._pp, self._ps, self._pst, actions.data_ptr(), lo.row0.data_ptr(), lo.posfeat.data_ptr(),
                                 rewards.data_ptr(), dones.data_ptr(), self._stats.data_ptr() if self._stats is not None else None,
                                 self.step_count, self._stream()),
            "fe_step_lazy",
        )
        info_dict = self.record_evaluation_metrics() if self.evaluate else {}
        return (lo, rewards, dones, info_dict)

    def step_host(self, actions_host: torch.Tensor, packed_dones: bool = False) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, dict]:
        """step() for a host-resident policy, through fe_step_host: `actions_host` is a CPU tensor
        (pinned for full speed); rewards and dones come back as pinned CPU tensors, already complete when the
        call returns (the pinned buffers are reused by the next step_host call; copy them to keep them);
        the observation stays in HBM.  Per step: 4N bytes host->device, 8N (f32) device->host.
        packed_dones=True (fe_step_host_packed): dones come back bit-packed — a uint8 tensor of 4*ceil(N/32) bytes, bit
        (i % 8) of byte i // 8 is env i's flag; `unpack_dones()` expands it — N/8 instead of 4N bytes on the wire."""
        if actions_host.device.type != "cpu" or actions_host.dtype != torch.float32 or not actions_host.is_contiguous():
            raise ValueError("actions_host must be a contiguous float32 CPU tensor")
        if actions_host.numel() != self.num_envs * self.num_acts:
            raise ValueError(f"actions must have {self.num_envs * self.num_acts} elements")
        if not hasattr(self, "_host_bufs"):
            self._host_bufs = (
                torch.empty(self.num_envs * self.num_acts, dtype=torch.float32, device=self._dev),
                torch.empty(self.num_envs, dtype=self.obs_dtype, device=self._dev),
                torch.empty(self.num_envs, dtype=torch.int32, device=self._dev),
                torch.empty(self.num_envs, dtype=self.obs_dtype).pin_memory(),
                torch.empty(self.num_envs, dtype=torch.int32).pin_memory(),
                torch.zeros(4 * ((self.num_envs + 31) // 32), dtype=torch.uint8).pin_memory(),
            )
            self._host_ptrs = tuple(t.data_ptr() for t in self._host_bufs)   # the buffers live as long as the env
        _, _, _, rewards, dones, done_bits = self._host_bufs
        p_a, p_r, p_d, p_rh, p_dh, p_bh = self._host_ptrs
        obs = self._new_obs()
        self.step_count += 1
        fn = self._L.fe_step_host_packed if packed_dones else self._L.fe_step_host
        out_dones = done_bits if packed_dones else dones
        _lib.check(
            fn(self._pp, self._ps, self._pst, actions_host.data_ptr(), p_a, obs.data_ptr(), p_r, p_d, p_rh,
               p_bh if packed_dones else p_dh, self._stats_ptr, self.step_count, self._stream()),
            "fe_step_host_packed" if packed_dones else "fe_step_host",
        )
        info_dict = self.record_evaluation_metrics() if self.evaluate else {}
        return (obs, rewards, out_dones, info_dict)

    def unpack_dones(self, done_bits: torch.Tensor) -> torch.Tensor:
        """(N,) int32 flags from the bit-packed form step_host(packed_dones=True) returns."""
        bits = np.unpackbits(done_bits.numpy(), bitorder="little")[: self.num_envs]
        return torch.from_numpy(bits.astype(np.int32))

    def host_bytes_per_step(self, packed_dones: bool = False) -> Tuple[int, int]:
        """(host->device, device->host) bytes one step_host() call moves over PCIe."""
        osz = 8 if self.obs_dtype == torch.float64 else 4
        dones = 4 * ((self.num_envs + 31) // 32) if packed_dones else 4 * self.num_envs
        return 4 * self.num_envs * self.num_acts, osz * self.num_envs + dones

    # ------------------------------------------------------------------ captured rollouts (8f-4) ----
    def _observe_into(self, obs: torch.Tensor) -> None:
        _lib.check(self._L.fe_observe(self._pp, self._ps, self._pst, obs.data_ptr(), self._stream()), "fe_observe")

    _STATE_FIELDS = ("_seg", "_ptr", "_cash", "_long", "_short", "_margin", "_terminated", "_ep_return", "_ep_len", "_stats")

    def state_snapshot(self) -> dict:
        """Copy of the per-env state (device tensors) + the step ordinal; load_state_snapshot() restores it in place."""
        snap = {k: getattr(self, k).clone() for k in self._STATE_FIELDS if getattr(self, k, None) is not None}
        snap["step_count"] = self.step_count
        return snap

    def load_state_snapshot(self, snap: dict) -> None:
        for k in self._STATE_FIELDS:
            if k in snap:
                getattr(self, k).copy_(snap[k])
        self.step_count = snap["step_count"]
        self._epoch += 1

    def capture_rollout(self, policy, num_steps: int) -> CapturedRollout:
        """Record `num_steps` x (policy -> step) in a CUDA graph; see CapturedRollout."""
        return CapturedRollout(self, policy, num_steps)

    def evaluate_policy(self, policy, steps_per_replay: int = 32, max_steps: Optional[int] = None,
                        rollout: Optional[CapturedRollout] = None) -> Dict:
        """The reference's evaluation loop (PPO_LSTM_testing_SPY.py:43-52: step until info has "returns") with the
        "all envs terminated" test (:531) read once per `steps_per_replay` graph-replayed steps instead of once per
        step.  Terminated envs collect zero reward (:527-528), so running past the end changes nothing: the returned
        {"returns": (N,) f32} equals what the per-step loop returns.  Metrics are reset afterwards (:532-534) — in
        place, so `rollout` (a CapturedRollout of this env, e.g. the one returned in the result) can be passed back
        in to evaluate the next checkpoint of the same policy object without capturing again."""
        if not self.evaluate:
            raise RuntimeError("evaluate_policy needs an env constructed with evaluate=True")
        roll = rollout if rollout is not None else self.capture_rollout(policy, steps_per_replay)
        if roll.env is not self:
            raise ValueError("rollout was captured on another env")
        steps = 0
        while max_steps is None or steps < max_steps:
            roll.replay(refresh_obs=(steps == 0))
            steps += roll.num_steps
            if int(self._stats[1].item()) >= self.num_envs:
                info = {"returns": self._ep_return.clone(), "steps": steps, "rollout": roll}
                self._terminated.zero_()
                self._ep_return.zero_()
                self._stats.zero_()
                return info
        return {"steps": steps, "rollout": roll}

    def record_evaluation_metrics(self) -> Dict:
        """:523-536 — the per-env bookkeeping ran inside the kernel; here only the "all terminated"
        test (:531), which like the reference's torch.all() costs one host read per step."""
        n_terminated = int(self._stats[1].item())
        if n_terminated >= self.num_envs:
            info_dict = {"returns": self._ep_return.clone()}   # the reference hands out the old tensor and rebinds (:532-534)
            self.reset_evaluation_metrics()
            return info_dict
        return {}

    # ------------------------------------------------------------------ extensions --------------
    def _launch_reset_all(self, redraw: bool) -> None:
        _lib.check(self._L.fe_reset_all(self._pp, self._ps, self._pst, self.step_count, int(redraw), self._stream()),
                   "fe_reset_all")

    def reset_all(self, redraw: Optional[bool] = None, lazy: bool = False):
        """Fresh episode for every env (the reset the ES loop expects, cf. isaac_gym_env.py:55-58)."""
        if redraw is None:
            redraw = self.random_reset == "all"
        self._epoch += 1
        self._launch_reset_all(redraw)
        if self._stats is not None:
            self._stats.zero_()
        return self.reset_lazy() if lazy else self.reset()

    def kernel_name(self) -> str:
        """Which kernel step() launches for this env's shape (diagnostics)."""
        return self._L.fe_step_kernel_name(self._pp, self._ps, self._pst).decode()

    def stats(self) -> Dict[str, torch.Tensor]:
        """Device-side episode statistics accumulated since the last clear (track_stats=True)."""
        if self._stats is None:
            raise RuntimeError("construct the env with track_stats=True")
        f = self._stats.view(torch.float64)
        return {"n_done": self._stats[0], "n_terminated": self._stats[1], "sum_len": self._stats[2],
                "sum_return": f[4], "sum_return_sq": f[5]}

    def clear_stats(self) -> None:
        if self._stats is not None:
            self._stats.zero_()
