"""The reference's OWN unit tests for the rebuilt classes, run against the drop-ins with nothing changed but the import
(tests/unit/test_time_series_env.py, tests/unit/test_PPO_buffer.py, tests/unit/test_parallel_mlp.py,
tests/integration/test_SPY_training.py of hmomin/FinEnvs).  They assert types and shapes only — the numeric pin is
in the other test files — but they are what a user of the reference would run first after switching."""
from typing import Dict, Tuple

import pytest
import torch

from parity_utils import load_trace

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def data_dir(tmp_path_factory):
    """finenvs/data/{IBM,OIH,SPY}/dummy.csv re-serialised from the golden fixture (the checkout is not on the GPU box)."""
    base = tmp_path_factory.mktemp("finenvs") / "root"
    zc = load_trace("dummy_csv.npz")
    for instr in ("IBM", "OIH", "SPY"):
        d = base / instr
        d.mkdir(parents=True)
        for key in ("dummy", "train"):     # the reference's SPY "train" file is not distributed: the dummy bars stand in
            with open(d / f"{key}.csv", "w") as f:
                for date, time, ohlc, vol in zip(zc[f"{instr}_csv_date"], zc[f"{instr}_csv_time"], zc[f"{instr}_csv_ohlc"],
                                                 zc[f"{instr}_csv_volume"]):
                    f.write(f"{date.decode()},{time.decode()},{float(ohlc[0])!r},{float(ohlc[1])!r},{float(ohlc[2])!r},"
                            f"{float(ohlc[3])!r},{vol}\n")
    mp = pytest.MonkeyPatch()
    mp.setenv("FINENVS_DATA_DIR", str(base))
    yield str(base)
    mp.undo()


def _step_helper(env):
    """tests/unit/test_time_series_env.py:24-36 verbatim."""
    num_envs = env.num_envs
    actions = torch.rand((num_envs, 1), device=env.device) * 2 - 1
    step_info: Tuple[torch.Tensor, torch.Tensor, torch.Tensor, Dict] = env.step(actions)
    assert isinstance(step_info, tuple)
    (next_states, rewards, dones, info) = step_info
    assert isinstance(next_states, torch.Tensor)
    assert isinstance(rewards, torch.Tensor)
    assert isinstance(dones, torch.Tensor)
    assert isinstance(info, dict)
    return next_states, rewards, dones


def test_time_series_env_suite(data_dir):
    """TestTimeSeriesEnv: default constructor (num_intervals = 390), reset, step, 1000 steps on the three fixtures."""
    from finenvs_b200.environments import TimeSeriesEnv

    envs = [TimeSeriesEnv("IBM", "dummy"), TimeSeriesEnv("OIH", "dummy"), TimeSeriesEnv("SPY", "dummy")]
    for env in envs:
        assert isinstance(env.reset(), torch.Tensor)
    for env in envs:
        _step_helper(env)
    for _ in range(1000):
        for env in envs:
            ns, r, d = _step_helper(env)
            # what the reference's outputs look like (time_series_env.py:277-296)
            assert ns.shape == (env.num_envs, 390, 5) and r.shape == d.shape == (env.num_envs,) and d.dtype == torch.int32
    for env in envs:
        ns, r, d = _step_helper(env)
        assert torch.isfinite(ns).all() and torch.isfinite(r).all()


def test_spy_training_suite(data_dir):
    """TestSPYTraining: TimeSeriesEnv("SPY", "train", evaluate=False), 1000 steps."""
    from finenvs_b200.environments import TimeSeriesEnv

    env = TimeSeriesEnv("SPY", "train", evaluate=False)
    assert isinstance(env.reset(), torch.Tensor)
    for _ in range(1001):
        _step_helper(env)


def test_ppo_buffer_suite():
    """TestBuffer, tests 1-7 in order."""
    from finenvs_b200.agents.PPO import Buffer

    num_envs, num_observations, num_actions, num_steps = 64, 24, 3, 4
    buffer = Buffer()

    def store_sample_data():
        states = torch.rand((num_envs, num_observations), device=buffer.device) * 2 - 1
        actions = torch.rand((num_envs, num_actions), device=buffer.device) * 2 - 1
        rewards = torch.ones((num_envs,), device=buffer.device)
        dones = torch.randint(0, 2, (num_envs,), device=buffer.device)
        log_probs = torch.rand((num_envs, num_actions), device=buffer.device) - 1
        values = torch.rand((num_envs, 1), device=buffer.device) * 2 - 1
        buffer.store(states, actions, rewards, dones, log_probs, values)

    for _ in range(num_steps):
        store_sample_data()
    assert buffer.size() == num_envs * num_steps
    last_values = torch.rand((num_envs, 1), device=buffer.device) * 2 - 1
    buffer.compute_returns_and_advantages(last_values)
    buffer.reshape()            # the reference's container is filled by store(); here by reshape() (time-major storage)
    rewards, returns, advantages = buffer.container["rewards"], buffer.container["returns"], buffer.container["advantages"]
    assert returns.shape == rewards.shape and advantages.shape == rewards.shape
    size = buffer.size()
    for item in buffer.container.values():
        assert item.shape[0] == size
    buffer.shuffle()
    batch_dict = buffer.get_batches()
    assert isinstance(batch_dict, dict)
    for key, value in batch_dict.items():
        assert isinstance(key, str) and isinstance(value, torch.Tensor)
    assert buffer.num_mini_batches == len(buffer.get_mini_batch_indices())


def test_parallel_mlp_suite():
    """TestParallelMLPNetworks, tests 1-4 in order."""
    from finenvs_b200.agents.networks import ParallelMLP

    num_envs, num_eval_envs, num_observations, num_actions = 8, 2, 24, 3
    network = ParallelMLP(num_envs, num_eval_envs, (num_observations, 5, 4, num_actions), noise_std_dev=0.02)
    states = torch.randn((num_envs, num_observations), device=network.device)
    fitnesses = torch.randn((num_envs,), device=network.device)
    network.perturb_parameters()
    outputs = network.forward(states)
    assert isinstance(outputs, torch.Tensor) and outputs.shape == (num_envs, num_actions)
    network.reconstruct_perturbations()
    network.update_parameters(fitnesses)
