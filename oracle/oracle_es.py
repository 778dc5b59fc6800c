"""CPU restatement of the reference's ES path (TEST INFRASTRUCTURE ONLY — never imported by the product).

Follows finenvs/agents/networks/parallel_mlp.py and finenvs/agents/ES/evo_agent.py, restated on the
representation the B200 build stores: UNIT perturbations eps per mirrored pair (the reference stores
base + sigma*eps per env, parallel_mlp.py:114-155).

  forward            parallel_mlp.py:84-103   per-env x @ W' + b', tanh, with W' = W +- sigma*eps (eval envs: W)
  update_parameters  parallel_mlp.py:176-275  mean over pairs of (f+ - f-) * eps  - l2 * theta, then Adam
  store / ranks      evo_agent.py:96-112, 154-186

Pinned against the reference itself: tests/golden/es_path.npz (tests/golden/make_golden_es.py) and live in
tests/test_oracle_es.py where the checkout exists.  float32 throughout, like the reference on its device.
"""
from __future__ import annotations

import numpy as np


def forward(weights, biases, eps_w, eps_b, sigma, num_eval, obs):
    """weights[l] (in,out) f32, biases[l] (1,out) f32, eps_w[l] (pairs,in,out), eps_b[l] (pairs,1,out) unit
    perturbations, obs (N,in0) f32 -> actions (N,out_last) f32 (no exploration noise)."""
    N = obs.shape[0]
    pairs = eps_w[0].shape[0]
    assert N == 2 * pairs + num_eval
    x = obs.astype(np.float32)[:, None, :]                                        # :87 unsqueeze(1)
    s = np.float32(sigma)
    for W, b, ew, eb in zip(weights, biases, eps_w, eps_b):
        zw = np.zeros((num_eval, *W.shape), np.float32)
        zb = np.zeros((num_eval, *b.shape), np.float32)
        pw = W[None] + np.concatenate([s * ew, -(s * ew), zw], 0)                  # :114-152
        pb = b[None] + np.concatenate([s * eb, -(s * eb), zb], 0)
        x = np.tanh(np.matmul(x, pw) + pb).astype(np.float32)                      # :91-95
    return x[:, 0, :]


class Adam:
    """parallel_mlp.py:66-82, 220-275"""

    def __init__(self, weights, biases, lr):
        self.t, self.lr, self.b1, self.b2 = 0, lr, 0.9, 0.999
        self.m_w = [np.zeros_like(w) for w in weights]
        self.m_b = [np.zeros_like(b) for b in biases]
        self.v_w = [np.zeros_like(w) for w in weights]
        self.v_b = [np.zeros_like(b) for b in biases]


def update_parameters(weights, biases, eps_w, eps_b, adam: Adam, fitnesses, num_eval, l2):
    """In-place update of weights / biases (lists of f32 arrays) from per-env fitnesses (N,) f32."""
    adam.t += 1
    pairs = eps_w[0].shape[0]
    f = fitnesses.astype(np.float32)
    diff = f[:pairs] - f[pairs: 2 * pairs]                                         # :178-185
    t, a, b1, b2 = adam.t, adam.lr, np.float32(adam.b1), np.float32(adam.b2)
    alpha_t = np.float32(np.sqrt(1 - adam.b2**t) / (1 - adam.b1**t) * a)           # :262
    for i, (W, b) in enumerate(zip(weights, biases)):
        gw = (diff[:, None, None] * eps_w[i]).astype(np.float32).mean(0, dtype=np.float32)   # :200-207 (sigma cancels)
        gb = (diff[:, None, None] * eps_b[i]).astype(np.float32).mean(0, dtype=np.float32)
        gw = gw - np.float32(l2) * W                                               # :208-209
        gb = gb - np.float32(l2) * b
        adam.m_w[i] = b1 * adam.m_w[i] + (np.float32(1) - b1) * gw                 # :246-261
        adam.m_b[i] = b1 * adam.m_b[i] + (np.float32(1) - b1) * gb
        adam.v_w[i] = b2 * adam.v_w[i] + (np.float32(1) - b2) * np.square(gw)
        adam.v_b[i] = b2 * adam.v_b[i] + (np.float32(1) - b2) * np.square(gb)
        W += alpha_t * (adam.m_w[i] / (np.sqrt(adam.v_w[i]) + np.float32(1e-8)))   # :263-274
        b += alpha_t * (adam.m_b[i] / (np.sqrt(adam.v_b[i]) + np.float32(1e-8)))


class Accounting:
    """evo_agent.py:36-40, 90-112, 154-186: running returns, finished-episode list, centred ranks."""

    def __init__(self, num_envs):
        self.N = num_envs
        self.cur_ret = np.zeros(num_envs, np.float32)
        self.cur_steps = np.zeros(num_envs, np.float32)
        self.total_timesteps = 0
        self.fin_ret = np.zeros(0, np.float32)
        self.dones = np.zeros(0, np.int64)

    def step_and_store(self, rewards, dones):
        self.cur_steps += 1                                                        # :93
        if rewards.dtype == np.float64:                                            # :99 f32 += f64: promoted, rounded once
            self.cur_ret = (self.cur_ret.astype(np.float64) + rewards).astype(np.float32)
        else:
            self.cur_ret = self.cur_ret + rewards
        idx = np.nonzero(dones)[0]                                                 # :100
        self.total_timesteps += float(self.cur_steps[idx].sum())                   # :102-103
        self.dones = np.concatenate([self.dones, idx])                             # :104
        self.fin_ret = np.concatenate([self.fin_ret, self.cur_ret[idx]])           # :105-107
        self.cur_ret[idx] = 0                                                      # :108-109
        self.cur_steps[idx] = 0
        return len(self.fin_ret), self.total_timesteps

    def final_ranks(self):
        order = np.argsort(self.fin_ret, kind="stable")                            # :165 (ties: unspecified in torch)
        ranks = np.empty(len(order), np.float32)
        ranks[order] = np.arange(len(order), dtype=np.float32)                     # :174-179
        centred = ranks / np.float32(len(ranks) - 1) - np.float32(0.5)             # :180-181
        out = np.zeros(self.N, np.float32)
        np.add.at(out, self.dones, centred)                                        # :184-186
        return centred, out

    def mean_returns(self):
        counts = np.bincount(self.dones, minlength=self.N)                         # :155-162
        num = np.zeros(self.N, np.float32)
        np.add.at(num, self.dones, self.fin_ret)
        den = counts.astype(np.float32)
        den[counts == 0] = 1.0
        return num / den
