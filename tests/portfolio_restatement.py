"""A SECOND, independent restatement of the multi-asset (A > 1) extension — TEST INFRASTRUCTURE ONLY.

The reference (hmomin/FinEnvs) is single-asset, so `oracle/fe_oracle.c: feo_step_multi` and the CUDA portfolio kernels
were twins by one author checked only against each other (missing #4).  This file states the same definition
(DESIGN.md §3 "Portfolio extension") a second time, in torch, ON TOP OF THE REAL REFERENCE OBJECTS:

* one unmodified reference `TimeSeriesEnv` per asset (built by its own loader from that asset's CSV) supplies every piece
  of index plumbing through its own methods — pointer / window advance (:281-282), `set_current_prices` (:323-342),
  `get_log_return_observations` (:437-445), `find_finished_environments` + `find_nan_spots` (:477-496),
  `reset_finished_environments` (:498-521) — and the elementwise helpers `get_share_changes_from_actions` (:298-302),
  `split_share_changes_by_sign` (:344-351);
* the bookkeeping below follows the reference's method bodies (:353-475, :288-289) line by line with every per-env
  quantity widened from (N, 1) to (N, A); dtype promotion and rounding are TORCH's (f32 tensors meeting f64 prices,
  in-place `+=` on an f32 account computing in f64 and rounding once) — nothing is hand-cast as in the C oracle;
* where the shared cash account meets per-asset amounts the definition says: one cash movement per phase — the
  xor-butterfly sum over assets of the per-asset amounts for the six phases whose amounts do not depend on cash, and a
  walk over the assets in index order with a running f64 balance for the two all-or-nothing entry phases.

With A = 1 the butterfly sum of one value is that value and the walk is one step: the class must then reproduce the real
reference's `step()` bit for bit (tests/test_portfolio_restatement.py checks that first, then A in {2, 5, 30} against
feo_step_multi).
"""
from __future__ import annotations

from typing import List

import torch


def butterfly_sum(x: torch.Tensor) -> torch.Tensor:
    """(N, A) -> (N, 1): the sum a 32-lane warp computes with five xor-shuffle rounds (lanes >= A hold 0)."""
    N, A = x.shape
    w = torch.zeros((N, 32), dtype=x.dtype)
    w[:, :A] = x
    lanes = torch.arange(32)
    for off in (16, 8, 4, 2, 1):
        w = w + w[:, lanes ^ off]
    return w[:, :1]


class PortfolioRestatement:
    def __init__(self, asset_envs: List, num_envs: int):
        """asset_envs: one REAL reference TimeSeriesEnv (evaluate=True: no redraw) per asset, same calendar."""
        self.e = asset_envs
        self.A = len(asset_envs)
        e0 = asset_envs[0]
        self.N = N = int(num_envs)
        D = e0.price_environments.shape[0]
        for e in asset_envs:                      # SURVEY App. C.4 widening: env i on day i mod D
            e.env_indices = torch.arange(N, dtype=torch.int64) % D
            e.num_envs = N
            e.env_pointers = torch.zeros((N,), dtype=torch.int64)
            e.env_spots = torch.arange(0, e.num_intervals).repeat(N, 1)
            e.cash = e.starting_balance * torch.ones((N, 1))
            e.long_shares = torch.zeros((N, 1))
            e.short_shares = torch.zeros((N, 1))
            e.margin = torch.zeros((N, 1))
            e.reset_evaluation_metrics()
            if hasattr(e, "current_close_prices"):
                del e.current_close_prices
        self.starting_balance = e0.starting_balance
        self.per_share_commission = e0.per_share_commission
        self.initial_margin_requirement = e0.initial_margin_requirement
        self.maintenance_margin_requirement = e0.maintenance_margin_requirement
        self.cash = self.starting_balance * torch.ones((N, 1))           # :264 one shared account
        self.long_shares = torch.zeros((N, self.A))                      # :266-269 per asset
        self.short_shares = torch.zeros((N, self.A))
        self.margin = torch.zeros((N, self.A))

    # ------------------------------------------------------------------ helpers
    def _prices(self):
        for e in self.e:
            e.set_current_prices()                                        # REAL :323-342
        cat = lambda name: torch.cat([getattr(e, name) for e in self.e], dim=1)   # noqa: E731
        self.O, self.H, self.L, self.C = (cat(f"current_{k}_prices") for k in ("open", "high", "low", "close"))

    def observe(self) -> torch.Tensor:
        """:423-435 per asset -> (N, W, 5A), asset-major inside a row: [lr0..lr3, posfeat] x A."""
        if not hasattr(self, "C"):
            self._prices()
        lr = torch.stack([e.get_log_return_observations() for e in self.e], dim=2)          # REAL :437-445 -> (N, W, A, 4)
        current_positions = (self.long_shares - self.short_shares) * self.C                # :428-430
        scaled_positions = current_positions / self.starting_balance                       # :431
        W = lr.shape[1]
        pf = scaled_positions.unsqueeze(1).repeat(1, W, 1).unsqueeze(-1)                   # :432-433
        return torch.cat([lr, pf], dim=3).reshape(self.N, W, 5 * self.A)                    # :434

    # ------------------------------------------------------------------ one step (:277-296)
    def step(self, actions: torch.Tensor):
        e0, A, N = self.e[0], self.A, self.N
        c = self.per_share_commission
        share_changes = e0.get_share_changes_from_actions(actions)        # REAL :298-302 (elementwise)
        for e in self.e:
            e.env_pointers += 1                                           # :281
            e.env_spots += 1                                              # :282
        # ---- determine_new_states :304-321
        commissions = torch.zeros((N, A))                                 # :305
        self._prices()
        pos, neg = e0.split_share_changes_by_sign(share_changes)          # REAL :344-351
        # sell_long_positions :353-361
        new_long_shares = torch.relu(self.long_shares + neg)
        sell_long_shares = self.long_shares - new_long_shares
        neg += sell_long_shares
        commissions += sell_long_shares * c                               # :363-365
        self.cash += butterfly_sum(sell_long_shares * (self.O - c))
        self.long_shares = new_long_shares
        # buy_back_short_positions :367-383
        new_short_shares = torch.relu(self.short_shares - pos)
        buy_back_shares = self.short_shares - new_short_shares
        pos -= buy_back_shares
        commissions += buy_back_shares * c
        self.cash -= butterfly_sum(buy_back_shares * (self.O + c))
        self.short_shares = new_short_shares
        new_margin = self.initial_margin_requirement * self.short_shares * self.O
        change_in_margin = new_margin - self.margin
        self.cash -= butterfly_sum(change_in_margin)
        self.margin = new_margin
        # disallow_illegal_long_trades + initiate_long_trades :385-399 — walk the assets with a running f64 balance
        running = self.cash.double()
        for a in range(A):
            new_long_positions = pos[:, a:a + 1] * (self.O[:, a:a + 1] + c)
            new_cash = running - new_long_positions
            illegal = new_cash < 0
            pos[:, a:a + 1][illegal] = 0
            running = torch.where(illegal, running, new_cash)
        self.cash = running.float()
        commissions += pos * c
        self.long_shares += pos
        # disallow_illegal_short_trades + initiate_short_trades :401-421 — same walk
        running = self.cash.double()
        for a in range(A):
            na = neg[:, a:a + 1]
            short_commission = -na * c
            new_short_positions = -na * self.O[:, a:a + 1]
            initial_margin_requirement = self.initial_margin_requirement * new_short_positions
            new_cash = running - initial_margin_requirement - short_commission
            illegal = new_cash < 0
            na[illegal] = 0
            # :413-418 recomputed after the veto, exactly as initiate_short_trades does
            short_commission = -na * c
            initial_margin_requirement = self.initial_margin_requirement * (-na * self.O[:, a:a + 1])
            running = torch.where(illegal, running, running - (initial_margin_requirement + short_commission))
        self.cash = running.float()
        commissions += -neg * c
        self.margin = self.margin + self.initial_margin_requirement * (-neg * self.O)
        self.short_shares += -neg
        new_states = self.observe()                                       # :321
        # ---- determine_immediate_rewards :447-457
        dones = self.cash < 0

        def maintenance_margin_check(prices):                             # :459-468
            nonlocal dones
            short_positions = self.short_shares * prices
            maintenance_margin_requirement = short_positions * (1 + self.maintenance_margin_requirement)
            margin_calls = torch.relu(maintenance_margin_requirement - self.margin)
            self.cash -= butterfly_sum(margin_calls)
            self.margin += margin_calls
            dones = torch.logical_or(dones, self.cash < 0)
            return -margin_calls

        rewards = maintenance_margin_check(self.H)
        short_positions = self.short_shares * self.L                      # margin_release :470-475
        margin_release = torch.relu(self.margin - short_positions * self.initial_margin_requirement)
        self.margin -= margin_release
        self.cash += butterfly_sum(margin_release)
        rewards += maintenance_margin_check(self.C)
        bankrupt = dones.view(N)
        self.long_shares[bankrupt] = 0                                    # :452-453
        self.short_shares[bankrupt] = 0
        price_change = self.C - self.O
        rewards += (self.long_shares - self.short_shares) * price_change
        rewards -= commissions
        rewards = butterfly_sum(rewards)                                  # per-asset rewards -> the env's reward
        # ---- episode end :477-496 through the real reference (same calendar for every asset)
        e0.dones = dones
        dones = e0.find_finished_environments()                          # REAL
        num_shares = butterfly_sum(self.short_shares + self.long_shares)  # :288
        rewards -= dones * num_shares * c                                 # :289
        # ---- reset_finished_environments :498-521: portfolio state here, pointers through the real reference
        done_rows = dones.view(N)
        self.cash[dones] = self.starting_balance
        self.margin[done_rows] = 0
        self.long_shares[done_rows] = 0
        self.short_shares[done_rows] = 0
        for e in self.e:
            e.reset_finished_environments(dones)                          # REAL (evaluate=True: no redraw)
        return new_states, rewards.squeeze(1), dones.squeeze(1).int()
