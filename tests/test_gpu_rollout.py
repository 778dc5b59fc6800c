"""PPO rollout storage on the GPU (finenvs_b200/agents/PPO/buffer.py, csrc/fe_rollout.cu) against the oracle,
the reference's golden vectors, and a list-based restatement of the reference's cat-on-store container."""
import numpy as np
import pytest
import torch

from test_oracle_rollout import CASES

pytestmark = pytest.mark.gpu


def _run_kernel(rewards, dones, values, last, gamma):
    from finenvs_b200 import _lib

    T, N = rewards.shape
    dev = "cuda:0"
    r, d, v, l = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (rewards, dones, values, last))
    ret = torch.empty((T, N), dtype=torch.float32, device=dev)
    adv = torch.empty((T, N), dtype=torch.float32, device=dev)
    _lib.check(_lib.lib().fe_returns_advantages(r.data_ptr(), int(rewards.dtype == np.float64), d.data_ptr(), v.data_ptr(),
                                                l.data_ptr(), N, T, gamma, ret.data_ptr(), adv.data_ptr(),
                                                torch.cuda.current_stream().cuda_stream), "fe_returns_advantages")
    return ret.cpu().numpy(), adv.cpu().numpy()


@pytest.mark.parametrize("name", sorted(CASES))
def test_returns_kernel_reproduces_reference_vectors(name):
    c = CASES[name]
    ret, adv = _run_kernel(c["rewards"], c["dones"], c["values"], c["last_values"], float(c["gamma"]))
    assert np.array_equal(ret.T, c["returns"]) and np.array_equal(adv.T, c["advantages"])


@pytest.mark.parametrize("N,T,rdt", [(100_003, 64, np.float32), (70_001, 33, np.float64), (1, 5, np.float32), (257, 1, np.float64)])
def test_returns_kernel_vs_oracle(N, T, rdt):
    from oracle import oracle_rollout as orl

    rng = np.random.default_rng(N + T)
    rewards = rng.normal(0, 2, (T, N)).astype(rdt)
    dones = (rng.uniform(0, 1, (T, N)) < 0.07).astype(np.int32)
    values = rng.normal(0, 1, (T, N)).astype(np.float32)
    last = rng.normal(0, 1, N).astype(np.float32)
    ret, adv = _run_kernel(rewards, dones, values, last, 0.99)
    r_ref, a_ref = orl.returns_and_advantages(rewards, dones, values, last, 0.99)
    assert np.array_equal(ret, r_ref) and np.array_equal(adv, a_ref)


def test_buffer_class_matches_reference_vectors_and_interface():
    """Feed the golden case through Buffer.store / prepare_training_data: same numbers as the reference's
    container (up to the documented time-major sample order), same shapes and dtypes, shuffle keeps rows paired."""
    from finenvs_b200.agents.PPO.buffer import Buffer

    c = CASES["f64_n64_t64"]
    T, N = c["rewards"].shape
    buf = Buffer(4, float(c["gamma"]), 0, capacity=16)  # forces two capacity doublings
    g = torch.Generator().manual_seed(0)
    states = torch.randn((T, N, 3, 5), generator=g, dtype=torch.float64)
    actions = torch.arange(T * N, dtype=torch.float32).view(T, N, 1)   # unique: identifies each sample after the shuffle
    for t in range(T):
        buf.store(states[t].cuda(), actions[t].cuda(), torch.from_numpy(c["rewards"][t]).cuda(),
                  torch.from_numpy(c["dones"][t]).cuda(), actions[t].cuda() * 2, torch.from_numpy(c["values"][t]).cuda().unsqueeze(-1))
        assert buf.size() == N * (t + 1)
    buf.prepare_training_data(torch.from_numpy(c["last_values"]).cuda().unsqueeze(-1))
    k = buf.container
    assert k["states"].shape == (T * N, 3, 5) and k["states"].dtype == torch.float64
    assert k["returns"].shape == (T * N, 1) and k["advantages"].dtype == torch.float32 and k["dones"].dtype == torch.int32
    assert np.array_equal(k["returns"].view(T, N).cpu().numpy().T, c["returns"])
    assert np.array_equal(k["advantages"].view(T, N).cpu().numpy().T, c["advantages"])
    assert torch.equal(k["states"].view(T, N, 3, 5).cpu(), states) and torch.equal(k["actions"].view(T, N, 1).cpu(), actions)
    before = {key: k[key].clone() for key in buf.batch_keys}
    batches = buf.get_batches()
    assert set(batches) == {"states", "actions", "log_probs", "advantages", "returns"}
    # one permutation for every key: find it from the actions and check the other keys follow it
    perm = batches["actions"].view(-1).long()
    assert sorted(perm.tolist()) == list(range(T * N)) and perm.tolist() != list(range(T * N))
    for key in buf.batch_keys:
        assert torch.equal(batches[key], before[key][perm]), key
    idx = buf.get_mini_batch_indices()
    assert len(idx) == 4 and all(len(i) == T * N // 4 for i in idx) and int(idx[-1][-1]) == T * N - 1
    buf.clear()
    assert buf.size() == 0 and all(v is None for v in buf.container.values())


def test_bound_env_writes_observations_in_place():
    """bind_env: the step kernel writes each observation straight into the slot the agent later stores (no copy
    except the first observation after clear()); contents equal an unbound twin env's observations."""
    from finenvs_b200.agents.PPO.buffer import Buffer
    from finenvs_b200.data import loader
    from finenvs_b200.environments import TimeSeriesEnv
    from parity_utils import gbm_ohlc

    W, bars, days, N, T = 8, 40, 30, 3000, 6
    rng = np.random.default_rng(2)
    prices = np.round(gbm_ohlc(rng, bars * days, 0.02), 4)
    seg_start, seg_len = loader.regular_segments(bars * days, bars, W)
    series = loader.stage_series(prices, seg_start, seg_len, W, "cuda:0", torch.float32)
    kw = dict(num_intervals=W, device_id=0, series=series, num_envs=N, seed=5, random_reset="all", random_offset=True)
    env, twin = TimeSeriesEnv("a", **kw), TimeSeriesEnv("b", **kw)
    buf = Buffer(2, 0.99, 0, capacity=T)
    buf.bind_env(env)
    states, twin_states = env.reset(), twin.reset()
    g = torch.Generator(device="cuda").manual_seed(1)
    for rollout in range(3):
        kept = []
        for t in range(T):
            actions = torch.rand((N, 1), generator=g, device="cuda") * 2 - 1
            nxt, rewards, dones, _ = env.step(actions)
            twin_nxt, r2, d2, _ = twin.step(actions)
            slot = buf._store["states"][t]
            in_place = states.data_ptr() == slot.data_ptr()
            assert in_place == (not (rollout > 0 and t == 0)), (rollout, t)   # only the carried-over observation is copied
            buf.store(states, actions, rewards, dones, actions, rewards.unsqueeze(-1))
            kept.append(twin_states)
            assert torch.equal(rewards, r2) and torch.equal(dones, d2)
            states, twin_states = nxt, twin_nxt
        buf.prepare_training_data(torch.zeros(N, 1, device="cuda"))
        got = buf.container["states"].view(T, N, W, 5)
        assert torch.equal(got, torch.stack(kept)) and torch.equal(states, twin_states)
        buf.clear()


def _twin_envs(N, W=8, bars=40, days=30, seed=5):
    from finenvs_b200.data import loader
    from finenvs_b200.environments import TimeSeriesEnv
    from parity_utils import gbm_ohlc

    rng = np.random.default_rng(2)
    prices = np.round(gbm_ohlc(rng, bars * days, 0.02), 4)
    seg_start, seg_len = loader.regular_segments(bars * days, bars, W)
    series = loader.stage_series(prices, seg_start, seg_len, W, "cuda:0", torch.float32)
    kw = dict(num_intervals=W, device_id=0, series=series, num_envs=N, seed=seed, random_reset="all", random_offset=True)
    return TimeSeriesEnv("a", **kw), TimeSeriesEnv("b", **kw)


def _rollouts_match_twin(env, twin, buf, states, twin_states, N, W, T, rollouts=3):
    g = torch.Generator(device="cuda").manual_seed(1)
    for rollout in range(rollouts):
        kept = []
        for t in range(T):
            actions = torch.rand((N, 1), generator=g, device="cuda") * 2 - 1
            nxt, rewards, dones, _ = env.step(actions)
            twin_nxt, r2, d2, _ = twin.step(actions)
            buf.store(states, actions, rewards, dones, actions, rewards.unsqueeze(-1))
            kept.append(twin_states)
            assert torch.equal(rewards, r2) and torch.equal(dones, d2)
            assert torch.equal(nxt, twin_nxt), (rollout, t)       # the observation in hand was not clobbered by the store
            states, twin_states = nxt, twin_nxt
        buf.prepare_training_data(torch.zeros(N, 1, device="cuda"))
        got = buf.container["states"].view(T, N, W, 5)
        assert torch.equal(got, torch.stack(kept)), rollout
        assert torch.equal(states, twin_states)
        buf.clear()


def test_bind_env_after_the_first_reset():
    """bind_env() used to assume env.reset() is the next env call; binding while the agent already holds an
    observation made the next step() write into the slot that store() then overwrote.  Any order must work."""
    from finenvs_b200.agents.PPO.buffer import Buffer

    N, W, T = 3000, 8, 5
    env, twin = _twin_envs(N, W)
    states, twin_states = env.reset(), twin.reset()     # reset FIRST (plain tensor) ...
    buf = Buffer(2, 0.99, 0, capacity=T)
    buf.bind_env(env)                                   # ... then bind
    _rollouts_match_twin(env, twin, buf, states, twin_states, N, W, T)


def test_bind_env_mid_rollout_and_capacity_growth():
    from finenvs_b200.agents.PPO.buffer import Buffer

    N, W, T = 1000, 8, 7
    env, twin = _twin_envs(N, W)
    states, twin_states = env.reset(), twin.reset()
    buf = Buffer(2, 0.99, 0, capacity=2)                # grows twice inside the first rollout
    g = torch.Generator(device="cuda").manual_seed(9)
    kept = []
    for t in range(T):
        if t == 3:
            buf.bind_env(env)                           # mid-rollout, agent holds an unbound observation
        actions = torch.rand((N, 1), generator=g, device="cuda") * 2 - 1
        nxt, rewards, dones, _ = env.step(actions)
        twin_nxt, _, _, _ = twin.step(actions)
        buf.store(states, actions, rewards, dones, actions, rewards.unsqueeze(-1))
        kept.append(twin_states)
        assert torch.equal(nxt, twin_nxt), t
        states, twin_states = nxt, twin_nxt
    buf.prepare_training_data(torch.zeros(N, 1, device="cuda"))
    assert torch.equal(buf.container["states"].view(T, N, W, 5), torch.stack(kept))


@pytest.mark.parametrize("N,W", [(3001, 3), (7, 390 // 6), (18949, 30)])   # tile / tile / pipe kernels
def test_bound_env_with_slots_that_are_not_16_byte_multiples(N, W):
    """slot k of the states storage started at k*N*W*5*4 bytes; with odd N (the reference's default is D+1
    envs) every other slot was misaligned for the kernels' 16-byte stores (FE_EALIGN).  Slot strides are padded now."""
    from finenvs_b200.agents.PPO.buffer import Buffer

    assert (N * W * 5 * 4) % 16 != 0
    T = 4
    env, twin = _twin_envs(N, W, bars=W + 25, days=12)
    buf = Buffer(2, 0.99, 0, capacity=T)
    buf.bind_env(env)
    states, twin_states = env.reset(), twin.reset()
    assert states.data_ptr() % 16 == 0
    _rollouts_match_twin(env, twin, buf, states, twin_states, N, W, T, rollouts=2)


def test_two_steps_without_a_store_do_not_alias():
    from finenvs_b200.agents.PPO.buffer import Buffer

    N, W = 2000, 8
    env, twin = _twin_envs(N, W)
    buf = Buffer(2, 0.99, 0, capacity=4)
    buf.bind_env(env)
    env.reset(), twin.reset()
    a = torch.zeros((N, 1), device="cuda")
    o1, _, _, _ = env.step(a)
    t1, _, _, _ = twin.step(a)
    o2, _, _, _ = env.step(a)                           # no store in between: must not overwrite o1
    t2, _, _, _ = twin.step(a)
    assert o1.data_ptr() != o2.data_ptr()
    assert torch.equal(o1, t1) and torch.equal(o2, t2)
