"""The rollout-return oracle (oracle/oracle_rollout.py) against vectors the reference's own
finenvs/agents/PPO/buffer.py produced (tests/golden/ppo_buffer.npz), and — where the checkout exists —
against the imported reference live."""
import os

import numpy as np
import pytest

from oracle import oracle_rollout as orl
from oracle import ref_harness as rh

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ppo_buffer.npz")


def golden_cases():
    with np.load(GOLD) as z:
        names = sorted({k.split(".")[0] for k in z.files})
        return {n: {k.split(".")[1]: z[k] for k in z.files if k.startswith(n + ".")} for n in names}


CASES = golden_cases()


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_reference_buffer_vectors(name):
    c = CASES[name]
    ret, adv = orl.returns_and_advantages(c["rewards"], c["dones"], c["values"], c["last_values"], float(c["gamma"]))
    assert ret.dtype == np.float32 and adv.dtype == np.float32
    assert np.array_equal(ret.T, c["returns"]) and np.array_equal(adv.T, c["advantages"])


@pytest.mark.skipif(not rh.available(), reason="reference checkout not present")
@pytest.mark.parametrize("rdt", ["float32", "float64"])
def test_oracle_vs_live_reference_buffer(rdt):
    import torch

    rh.ref_module()
    from finenvs.agents.PPO.buffer import Buffer

    rng = np.random.default_rng(3)
    N, T = 50, 37
    rewards = rng.normal(0, 2, (T, N)).astype(rdt)
    dones = (rng.uniform(0, 1, (T, N)) < 0.1).astype(np.int32)
    values = rng.normal(0, 1, (T, N)).astype(np.float32)
    last = rng.normal(0, 1, N).astype(np.float32)
    buf = Buffer(4, 0.97, -1)
    for t in range(T):
        buf.store(torch.zeros(N, 1, 5), torch.zeros(N, 1), torch.from_numpy(rewards[t]), torch.from_numpy(dones[t]),
                  torch.zeros(N, 1), torch.from_numpy(values[t]).unsqueeze(-1))
    buf.compute_returns_and_advantages(torch.from_numpy(last).unsqueeze(-1))
    ret, adv = orl.returns_and_advantages(rewards, dones, values, last, 0.97)
    assert np.array_equal(ret.T, buf.container["returns"].squeeze(-1).numpy())
    assert np.array_equal(adv.T, buf.container["advantages"].squeeze(-1).numpy())
