"""Host-side multi-GPU logic on CPU: world_size-2 gloo process group (the N>1 path of bench.py uses
the same functions over NCCL)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from finenvs_b200 import parallel as par

    r, w, _ = par.init_distributed("gloo")
    assert (r, w) == (rank, world)
    base, count = par.shard_bounds(total, rank, world)
    # every rank owns a contiguous block; together they tile [0, total)
    spans = [par.shard_bounds(total, k, world) for k in range(world)]
    assert spans[0][0] == 0 and all(spans[k][0] + spans[k][1] == spans[k + 1][0] for k in range(world - 1))
    assert spans[-1][0] + spans[-1][1] == total
    # fitness all-gather restores global env order; centred ranks equal the single-process result
    g = torch.Generator().manual_seed(0)
    fitness = torch.randn(total, generator=g)
    gathered = par.all_gather_fitness(fitness[base:base + count].clone(), total)
    assert torch.equal(gathered, fitness)
    local_ranks = par.global_centered_ranks(fitness[base:base + count].clone(), total)
    assert torch.equal(local_ranks, par.centered_ranks(fitness)[base:base + count])
    # episode statistics: sum of per-rank vectors
    vec = torch.tensor([3.0 + rank, 0.0, 30.0 * (rank + 1), 1.5 * (rank + 1), 2.0], dtype=torch.float64)
    st = par.all_reduce_episode_stats(vec)
    assert st["episodes"] == 7.0 and st["mean_length"] == 90.0 / 7.0
    assert abs(st["mean_return"] - 4.5 / 7.0) < 1e-12
    # evaluate-mode flag: all-reduce(min)
    assert par.all_terminated(torch.tensor(rank == 0)) is False
    assert par.all_terminated(torch.tensor(True)) is True
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [10, 11])
def test_world_size_2_gloo(total):
    mp.spawn(_worker, args=(2, _free_port(), total), nprocs=2, join=True)


def test_shard_bounds_single_process():
    from finenvs_b200 import parallel as par

    assert par.shard_bounds(8, 0, 1) == (0, 8)
    assert [par.shard_bounds(10, r, 4) for r in range(4)] == [(0, 3), (3, 3), (6, 2), (8, 2)]
    with pytest.raises(ValueError):
        par.shard_bounds(2, 0, 4)
    assert par.all_reduce_episode_stats(torch.zeros(5, dtype=torch.float64))["episodes"] == 0.0
