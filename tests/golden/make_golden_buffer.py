"""Generates tests/golden/ppo_buffer.npz by running the REAL reference rollout buffer
(finenvs/agents/PPO/buffer.py) in this container.

Run:  python tests/golden/make_golden_buffer.py        (needs /root/reference; CPU only)

Cases: rewards f32 (what finenvs_b200's env returns by default) and f64 (what the reference env returns,
time_series_env.py:296), several (N, T), dones dense enough that most envs cross an episode boundary.
Stored per case: the inputs in the order store() received them (time-major, (T, N)) and the reference's
`returns` / `advantages` (buffer.py:80-100) reshaped back to (N, T).
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402


def main():
    import torch

    rh.ref_module()  # puts the shim + the reference checkout on sys.path
    from finenvs.agents.PPO.buffer import Buffer

    out = {}
    cases = [("f32_n7_t9", 7, 9, torch.float32, 0.99, 0.3), ("f64_n7_t9", 7, 9, torch.float64, 0.99, 0.3),
             ("f32_n64_t64", 64, 64, torch.float32, 0.99, 0.05), ("f64_n64_t64", 64, 64, torch.float64, 0.9, 0.05),
             ("f32_n3_t1", 3, 1, torch.float32, 0.99, 0.5), ("f64_n3_t1", 3, 1, torch.float64, 0.99, 0.5),
             ("f32_n33_t200", 33, 200, torch.float32, 0.999, 0.01)]
    for name, N, T, rdt, gamma, pdone in cases:
        g = torch.Generator().manual_seed(N * 1000 + T + (1 if rdt == torch.float64 else 0))
        buf = Buffer(4, gamma, -1)
        rewards = (torch.randn((T, N), generator=g, dtype=torch.float64) * 3).to(rdt)
        dones = (torch.rand((T, N), generator=g) < pdone).int()
        values = torch.randn((T, N, 1), generator=g)
        last_values = torch.randn((N, 1), generator=g)
        for t in range(T):
            buf.store(torch.zeros(N, 2, 5), torch.zeros(N, 1), rewards[t], dones[t], torch.zeros(N, 1), values[t])
        buf.compute_returns_and_advantages(last_values)
        ret, adv = buf.container["returns"], buf.container["advantages"]
        assert ret.shape == (N, T, 1) and ret.dtype == torch.float32 and adv.dtype == torch.float32
        out[f"{name}.rewards"] = rewards.numpy()
        out[f"{name}.dones"] = dones.numpy()
        out[f"{name}.values"] = values.squeeze(-1).numpy()
        out[f"{name}.last_values"] = last_values.squeeze(-1).numpy()
        out[f"{name}.gamma"] = np.float64(gamma)
        out[f"{name}.returns"] = ret.squeeze(-1).numpy()       # (N, T)
        out[f"{name}.advantages"] = adv.squeeze(-1).numpy()
    np.savez_compressed(os.path.join(HERE, "ppo_buffer.npz"), **out)
    print("wrote ppo_buffer.npz:", sorted({k.split('.')[0] for k in out}))


if __name__ == "__main__":
    main()
