"""EvoAgent — drop-in for the reference's finenvs/agents/ES/evo_agent.py:11-191 on top of the B200 ParallelMLP.

Same constructor and methods (`step`, `store`, `train`, `log_progress`, `compute_mean_returns`,
`perform_rank_transformation`, ...), so the reference's training loop (examples/isaac_gym/ES_MLP_Isaac_Gym.py:30-38)
runs as written.  Differences:

* the episode accounting of `store` (:96-112: nonzero + two cats + `.item()` per step) is one launch of
  `fe_es_store`: per-env running return / step count and an append-only device list of finished episodes.
  `store` still returns the reference's `(num_finished, total_timesteps)` — as lazy counters that read the device
  only when the caller compares or converts them; `store_async` returns nothing and never synchronises;
* `step` also accepts the env's LazyObs handle (no observation tensor exists on that path);
* sharded populations (one process per GPU): finished episodes are all-gathered so the centred-rank transform
  (:164-186) is ONE global sort, identical on every rank; the gradient sum is all-reduced inside ParallelMLP.
"""
from __future__ import annotations

import csv
import os
from datetime import datetime
from time import time
from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

from ... import _lib
from ...base_object import BaseObject
from ...device_utils import require_cuda_device
from ..networks.parallel_mlp import ParallelMLP


class _LazyCount:
    """A device counter that behaves like the int the reference returns; reads the device when used."""

    def __init__(self, tensor: torch.Tensor, index: int):
        self._t, self._i = tensor, index

    def __int__(self) -> int:
        return int(self._t[self._i].item())

    __index__ = __int__

    def __float__(self) -> float:
        return float(int(self))

    def __ge__(self, o): return int(self) >= o
    def __gt__(self, o): return int(self) > o
    def __le__(self, o): return int(self) <= o
    def __lt__(self, o): return int(self) < o
    def __eq__(self, o): return int(self) == o
    def __repr__(self): return str(int(self))


class EvoAgent(BaseObject):
    def __init__(
        self,
        env_args: Dict,
        hidden_dims: Tuple[int] = (256, 256),
        learning_rate: float = 0.01,
        noise_std_dev: float = 0.02,
        l2_coefficient: float = 0.005,
        write_to_csv: bool = True,
        device_id: int = 0,
        *,
        seed: Optional[int] = None,
        max_finished: Optional[int] = None,
        env_id_base: int = 0,
        total_envs: Optional[int] = None,
        pair_id_base: int = 0,
        total_pairs: Optional[int] = None,
        group=None,
    ):
        """Reference arguments: :13-22.  Extensions: `seed` (perturbation stream key), `max_finished` (capacity of the
        finished-episode list per generation, default max(8 x num_envs, 65536); 20 bytes per entry), and the shard description
        (`env_id_base`, `total_envs`, `pair_id_base`, `total_pairs`, `group`) for multi-GPU populations."""
        self.set_env_params(env_args)
        self.device = require_cuda_device(device_id)
        self._dev = torch.device(self.device)
        self.learning_rate = learning_rate
        self.noise_std_dev = noise_std_dev
        self.network_shape = (self.num_observations, *hidden_dims, self.num_actions)
        self.env_id_base = int(env_id_base)
        self.total_envs = int(self.num_envs if total_envs is None else total_envs)
        self.group = group
        self.network = ParallelMLP(
            self.num_envs, self.num_eval_envs, self.network_shape, learning_rate=learning_rate,
            noise_std_dev=noise_std_dev, l2_coefficient=l2_coefficient, device_id=device_id, seed=seed,
            pair_id_base=pair_id_base, total_pairs=total_pairs, env_id_base=env_id_base, group=group,
        )
        self.network.perturb_parameters()
        self._L = _lib.lib()
        self.max_finished = int(max_finished if max_finished is not None else max(8 * self.num_envs, 1 << 16))
        self._fin_key = torch.empty(self.max_finished, dtype=torch.int64, device=self._dev)
        self._fin_env = torch.empty(self.max_finished, dtype=torch.int64, device=self._dev)
        self._fin_ret = torch.empty(self.max_finished, dtype=torch.float32, device=self._dev)
        self._counters = torch.zeros(2, dtype=torch.int64, device=self._dev)   # [finished episodes, total timesteps]
        self._store_calls = 0
        self._reset_accumulators()
        self.write_to_csv = write_to_csv
        if write_to_csv:
            self.create_progress_log()

    def _reset_accumulators(self) -> None:
        """:148-152"""
        self.current_returns = torch.zeros((self.num_envs,), device=self._dev)
        self.current_timesteps = torch.zeros((self.num_envs,), device=self._dev)
        self._counters.zero_()

    # ------------------------------------------------------------------ env description, progress log (:48-88)
    _ENV_KEYS = {
        "env_name": "label used in log file names",
        "num_envs": "size of the whole population held by this agent (training + evaluation envs)",
        "num_eval_envs": "how many of them, at the end, run the unperturbed parameters",
        "num_observations": "length of one env's observation vector",
        "num_actions": "length of one env's action vector",
    }
    _LOG_COLUMNS = ("unix_time", "num_episodes", "mean_eval_return", "std_dev_eval_return", "mean_training_return",
                    "std_dev_training_return", "L2_norm")   # the reference's CSV schema (:68-76): a file format, kept

    def set_env_params(self, env_args: Dict) -> None:
        missing = [k for k in self._ENV_KEYS if k not in env_args]
        if missing:
            wanted = "; ".join(f"{k} ({why})" for k, why in self._ENV_KEYS.items())
            raise Exception(f"env_args lacks {missing}. EvoAgent reads: {wanted}")
        self.env_name = env_args["env_name"]
        self.num_envs = int(env_args["num_envs"])
        self.num_eval_envs = int(env_args["num_eval_envs"])
        self.num_observations = int(env_args["num_observations"])
        self.num_actions = int(env_args["num_actions"])
        self.num_training_envs = self.num_envs - self.num_eval_envs

    def create_progress_log(self) -> None:
        """One CSV per run under ./trials, named <env>_Evo_<timestamp>.csv like the reference's (:62-67)."""
        folder = os.path.join(os.getcwd(), "trials")
        os.makedirs(folder, exist_ok=True)
        stamp = datetime.now().strftime("%Y-%m-%d_%H-%M-%S")
        self.csv_name = os.path.join(folder, f"{self.env_name}_Evo_{stamp}.csv")
        self.log = open(self.csv_name, "a", newline="")
        self.writer = csv.writer(self.log)
        self.writer.writerow(self._LOG_COLUMNS)

    # ------------------------------------------------------------------ rollout (:90-112)
    def step(self, states) -> torch.Tensor:
        """:90-94 (the per-env step counter is advanced inside store, which always follows)."""
        return self.network.forward(states)

    def store_async(self, rewards: torch.Tensor, dones: torch.Tensor) -> None:
        """:96-112 without host synchronisation."""
        if rewards.dtype not in (torch.float32, torch.float64) or dones.dtype != torch.int32:
            raise TypeError("rewards must be float32/float64 and dones int32 (what the env returns)")
        self._store_calls += 1
        _lib.check(
            self._L.fe_es_store(rewards.data_ptr(), int(rewards.dtype == torch.float64), dones.data_ptr(), self.num_envs,
                                self.env_id_base, self.total_envs, self._store_calls, self.current_returns.data_ptr(),
                                self.current_timesteps.data_ptr(), self.max_finished, self._counters.data_ptr(),
                                self._fin_key.data_ptr(), self._fin_env.data_ptr(), self._fin_ret.data_ptr(),
                                torch.cuda.current_stream(self._dev).cuda_stream),
            "fe_es_store",
        )

    def store(self, rewards: torch.Tensor, dones: torch.Tensor) -> Tuple[int, int]:
        self.store_async(rewards, dones)
        return (_LazyCount(self._counters, 0), _LazyCount(self._counters, 1))

    # the reference's attributes, derived from the device list (sorted into the reference's order: by step, then env)
    def _finished(self) -> Tuple[torch.Tensor, torch.Tensor]:
        n = int(self._counters[0].item())
        if n > self.max_finished:
            raise RuntimeError(f"{n} episodes finished this generation but max_finished={self.max_finished}")
        order = torch.argsort(self._fin_key[:n])
        return self._fin_env[:n][order], self._fin_ret[:n][order]

    @property
    def dones(self) -> torch.Tensor:
        return self._finished()[0]

    @property
    def finished_returns(self) -> torch.Tensor:
        return self._finished()[1]

    @property
    def total_timesteps(self) -> int:
        return int(self._counters[1].item())

    # ------------------------------------------------------------------ training (:114-191)
    def _world(self) -> int:
        return dist.get_world_size(self.group) if dist.is_initialized() else 1

    def _gather_finished(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """(global env ids, returns) of every finished episode of the generation, same order on every rank."""
        env, ret = self._finished()
        env = env + self.env_id_base
        world = self._world()
        if world == 1:
            return env, ret
        n = torch.tensor([env.numel()], dtype=torch.int64, device=self._dev)
        counts = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(counts, n, group=self.group)
        counts = [int(c.item()) for c in counts]
        width = max(max(counts), 1)
        pe = torch.zeros(width, dtype=torch.int64, device=self._dev); pe[: env.numel()] = env
        pr = torch.zeros(width, dtype=torch.float32, device=self._dev); pr[: ret.numel()] = ret
        ge = torch.empty(world * width, dtype=torch.int64, device=self._dev)
        gr = torch.empty(world * width, dtype=torch.float32, device=self._dev)
        dist.all_gather_into_tensor(ge, pe, group=self.group)
        dist.all_gather_into_tensor(gr, pr, group=self.group)
        keep = torch.cat([torch.arange(r * width, r * width + c, device=self._dev) for r, c in enumerate(counts)])
        return ge[keep], gr[keep]

    def train(self) -> float:
        """One ES generation update (:164-171): ranks of the finished episodes -> parameter update -> statistics ->
        fresh perturbations for the next generation.  Returns the mean evaluation return like the reference."""
        net = self.network
        net.reconstruct_perturbations()          # a no-op here (perturbations are a pure function of their counters)
        self.perform_rank_transformation()
        net.update_parameters(self.final_ranks)
        mean_eval = self.log_progress()
        net.perturb_parameters()
        return mean_eval

    def perform_rank_transformation(self) -> None:
        """:173-191 — ONE sort over the finished episodes of ALL ranks, so every rank derives the same ranks."""
        self._g_env, self._g_ret = self._gather_finished()
        self.compute_centered_ranks(self._g_ret.argsort())
        self.compute_final_ranks()

    def compute_centered_ranks(self, sort_indices: torch.Tensor) -> None:
        """Position of every episode in the sorted order, mapped linearly onto [-0.5, 0.5] (:176-181)."""
        n = sort_indices.numel()
        position = torch.empty(n, dtype=torch.float32, device=self._dev)
        position[sort_indices] = torch.arange(n, dtype=torch.float32, device=self._dev)
        self.centered_ranks = position / (n - 1) - 0.5

    def compute_final_ranks(self) -> None:
        """:183-186, restricted to the envs this rank holds: an env's weight is the sum of its episodes' ranks."""
        lo, hi = self.env_id_base, self.env_id_base + self.num_envs
        mine = (self._g_env >= lo) & (self._g_env < hi)
        self.final_ranks = torch.zeros(self.num_envs, dtype=torch.float32, device=self._dev)
        self.final_ranks.index_add_(0, self._g_env[mine] - lo, self.centered_ranks[mine])

    def compute_mean_returns(self) -> None:
        """Per-env mean over the episodes it finished this generation; envs without one report 0 (:154-162)."""
        env_of_episode, episode_return = self._finished()
        total = torch.zeros(self.num_envs, dtype=torch.float32, device=self._dev)
        total.index_add_(0, env_of_episode, episode_return)
        episodes = torch.bincount(env_of_episode, minlength=self.num_envs)
        self.mean_returns = total / episodes.clamp(min=1).to(torch.float32)

    def log_progress(self) -> float:
        """:114-152 (statistics of this rank's envs; rank 0 of a sharded run prints)."""
        self.compute_mean_returns()
        train_part = self.mean_returns[: self.num_training_envs]
        eval_part = self.mean_returns[self.num_training_envs:]
        nan = float("nan")
        row = {
            "unix_time": time(),
            "num_episodes": int(self._counters[0].item()),
            "mean_eval_return": eval_part.mean().item() if self.num_eval_envs else nan,
            "std_dev_eval_return": eval_part.std().item() if self.num_eval_envs > 1 else nan,
            "mean_training_return": train_part.mean().item(),
            "std_dev_training_return": train_part.std().item(),
            "L2_norm": self.network.get_l2_norm(),
        }
        if self.write_to_csv:
            self.writer.writerow([row[c] for c in self._LOG_COLUMNS])
            self.log.flush()
        if not dist.is_initialized() or dist.get_rank(self.group) == 0:
            print(f"[ES] episodes {row['num_episodes']}  env-steps {self.total_timesteps}  "
                  f"eval {row['mean_eval_return']:.2f} +- {row['std_dev_eval_return']:.2f}  "
                  f"train {row['mean_training_return']:.2f} +- {row['std_dev_training_return']:.2f}  "
                  f"|theta| {row['L2_norm']:.2f}")
        del self.centered_ranks, self.final_ranks, self.mean_returns      # per-generation scratch, as in the reference
        self._reset_accumulators()
        return row["mean_eval_return"]
