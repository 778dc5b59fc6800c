"""Generates tests/golden/es_path.npz by running the REAL reference ES classes
(finenvs/agents/networks/parallel_mlp.py, finenvs/agents/ES/evo_agent.py) in this container.

Run:  python tests/golden/make_golden_es.py        (needs /root/reference; CPU only)

The reference draws its perturbations with torch.normal inside perturb_parameters; to compare representations
the script sets `perturbed_weights = base +- sigma * eps` from its OWN fp16-representable unit perturbations eps
(exactly what the B200 build stores) and disables the exploration noise (add_action_noise) so forward is
deterministic.  Everything else — forward, update_parameters + Adam, EvoAgent.store, the rank transform,
compute_mean_returns — is the reference's code, unmodified.
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

    rh.ref_module()
    from finenvs.agents.ES.evo_agent import EvoAgent
    from finenvs.agents.networks.parallel_mlp import ParallelMLP

    out = {}
    cases = [("mlp_20_8_1", (20, 8, 1), 64, 4, 0.02), ("mlp_300_8_1", (300, 8, 1), 40, 0, 0.05),
             ("mlp_15_12_5_3", (15, 12, 5, 3), 30, 2, 0.1), ("lin_7_2", (7, 2), 10, 2, 0.02)]
    for name, shape, N, E, sigma in cases:
        torch.manual_seed(len(name) * 131 + N)
        net = ParallelMLP(N, E, shape, learning_rate=0.01, noise_std_dev=sigma, l2_coefficient=0.005, device_id=-1)
        net.add_action_noise = lambda actions, std: None        # deterministic forward
        pairs = (N - E) // 2
        g = torch.Generator().manual_seed(N)
        eps_w = [torch.randn((pairs, *w.shape), generator=g).half().float() for w in net.weight_layers]
        eps_b = [torch.randn((pairs, *b.shape), generator=g).half().float() for b in net.bias_layers]

        def set_perturbed():
            net.perturbed_weights, net.perturbed_biases = [], []
            for w, b, ew, eb in zip(net.weight_layers, net.bias_layers, eps_w, eps_b):
                net.perturbed_weights.append(w.repeat((N, 1, 1)) + torch.cat([sigma * ew, -(sigma * ew), torch.zeros((E, *w.shape))], 0))
                net.perturbed_biases.append(b.repeat((N, 1, 1)) + torch.cat([sigma * eb, -(sigma * eb), torch.zeros((E, *b.shape))], 0))

        for i, (w, b) in enumerate(zip(net.weight_layers, net.bias_layers)):
            out[f"{name}.w{i}"], out[f"{name}.b{i}"] = w.numpy().copy(), b.numpy().copy()
            out[f"{name}.eps_w{i}"], out[f"{name}.eps_b{i}"] = eps_w[i].numpy(), eps_b[i].numpy()
        out[f"{name}.shape"], out[f"{name}.num_eval"], out[f"{name}.sigma"] = np.array(shape), np.int64(E), np.float64(sigma)
        obs = torch.randn((N, shape[0]), generator=g) * 0.7
        out[f"{name}.obs"] = obs.numpy()
        set_perturbed()
        out[f"{name}.actions"] = net.forward(obs).numpy().copy()
        for u in range(2):                                       # two consecutive updates (Adam state carries over)
            fit = torch.randn(N, generator=g)
            out[f"{name}.fitness{u}"] = fit.numpy()
            if u:
                set_perturbed()
            net.reconstruct_perturbations()
            net.update_parameters(fit)
            for i, (w, b) in enumerate(zip(net.weight_layers, net.bias_layers)):
                out[f"{name}.w{i}_after{u}"], out[f"{name}.b{i}_after{u}"] = w.numpy().copy(), b.numpy().copy()
        out[f"{name}.l2_norm"] = np.float64(net.get_l2_norm())

    # EvoAgent accounting + rank transform (network untouched: only store / rank / mean paths are exercised)
    for name, N, E, T, rdt in [("acct_f32", 50, 4, 40, torch.float32), ("acct_f64", 31, 1, 25, torch.float64)]:
        torch.manual_seed(7)
        agent = EvoAgent({"env_name": "x", "num_envs": N, "num_eval_envs": E, "num_observations": 6, "num_actions": 1},
                         hidden_dims=(4,), write_to_csv=False, device_id=-1)
        g = torch.Generator().manual_seed(N + T)
        rewards = (torch.randn((T, N), generator=g, dtype=torch.float64)).to(rdt)
        dones = (torch.rand((T, N), generator=g) < 0.15).int()
        counts = []
        for t in range(T):
            agent.current_timesteps += 1                         # what EvoAgent.step does (:93)
            counts.append(agent.store(rewards[t], dones[t]))
        agent.perform_rank_transformation()
        agent.compute_mean_returns()
        out[f"{name}.rewards"], out[f"{name}.dones"] = rewards.numpy(), dones.numpy()
        out[f"{name}.num_eval"] = np.int64(E)
        out[f"{name}.counts"] = np.array(counts, dtype=np.float64)
        out[f"{name}.finished_returns"] = agent.finished_returns.numpy()
        out[f"{name}.done_envs"] = agent.dones.numpy()
        out[f"{name}.centered_ranks"] = agent.centered_ranks.numpy()
        out[f"{name}.final_ranks"] = agent.final_ranks.numpy()
        out[f"{name}.mean_returns"] = agent.mean_returns.numpy()
        out[f"{name}.current_returns"] = agent.current_returns.numpy()
    np.savez_compressed(os.path.join(HERE, "es_path.npz"), **out)
    print("wrote es_path.npz:", sorted({k.split('.')[0] for k in out}))


if __name__ == "__main__":
    main()
