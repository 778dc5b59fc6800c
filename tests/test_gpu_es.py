"""ES rollout path on the GPU (csrc/fe_es.cu, finenvs_b200/agents/{networks,ES}) against the oracle and the vectors
the reference's own ParallelMLP / EvoAgent produced (tests/golden/es_path.npz).

Floating-point kernel: the forward pass sums f32 dot products in a different order than torch.matmul, so the tolerance
is rtol 1e-5 (the north star's) + atol 2e-6 on tanh outputs in (-1, 1); everything integer (episode lists, counters,
pointers, dones) is exact."""
import numpy as np
import pytest
import torch

from test_oracle_es import ACCTS, CASES, NETS, net_arrays

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-5, 2e-6


def _series(W, days=40, bars=50, sigma=0.02, seed=3):
    from finenvs_b200.data import loader
    from parity_utils import gbm_ohlc

    rng = np.random.default_rng(seed)
    prices = np.round(gbm_ohlc(rng, bars * days, sigma), 4)
    seg_start, seg_len = loader.regular_segments(bars * days, bars, W)
    return loader.stage_series(prices, seg_start, seg_len, W, "cuda:0", torch.float32)


def _env(series, N, **kw):
    from finenvs_b200.environments import TimeSeriesEnv

    return TimeSeriesEnv("es", num_intervals=series.window, device_id=0, series=series, num_envs=N, seed=5,
                         random_reset="all", random_offset=True, **kw)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_lazy_step_equals_step(dtype):
    from finenvs_b200.data import loader
    from parity_utils import gbm_ohlc

    W, N = 12, 5001
    rng = np.random.default_rng(1)
    prices = np.round(gbm_ohlc(rng, 2000, 0.05), 4)
    seg_start, seg_len = loader.regular_segments(2000, 40, W)
    series = loader.stage_series(prices, seg_start, seg_len, W, "cuda:0", dtype)
    kw = dict(num_intervals=W, device_id=0, series=series, num_envs=N, seed=5, random_reset="all", random_offset=True,
              obs_dtype=dtype, track_stats=True)
    from finenvs_b200.environments import TimeSeriesEnv

    a, b = TimeSeriesEnv("a", **kw), TimeSeriesEnv("b", **kw)
    assert torch.equal(a.reset(), b.reset_lazy().materialize())
    g = torch.Generator(device="cuda").manual_seed(0)
    for t in range(80):
        act = torch.rand((N, 1), generator=g, device="cuda") * 2 - 1
        o, r, d, _ = a.step(act)
        lo, r2, d2, _ = b.step_lazy(act)
        assert torch.equal(r, r2) and torch.equal(d, d2)
        assert torch.equal(o, lo.materialize()), t
        for k in ("_seg", "_ptr", "_cash", "_long", "_short", "_margin"):
            assert torch.equal(getattr(a, k), getattr(b, k)), k
    assert int(a.stats()["n_done"]) == int(b.stats()["n_done"]) > 0


def _net_from_case(c, **kw):
    from finenvs_b200.agents.networks import ParallelMLP

    W, B, EW, EB = net_arrays(c)
    E = int(c["num_eval"])
    N = 2 * EW[0].shape[0] + E
    net = ParallelMLP(N, E, tuple(int(x) for x in c["shape"]), learning_rate=0.01, noise_std_dev=float(c["sigma"]),
                      l2_coefficient=0.005, device_id=0, seed=1, action_noise_std=0.0, **kw)
    for i in range(len(W)):
        net.weight_layers[i].copy_(torch.from_numpy(W[i]))
        net.bias_layers[i].copy_(torch.from_numpy(B[i]))
    net._theta_dirty = True
    net.set_perturbations([torch.from_numpy(e) for e in EW], [torch.from_numpy(e) for e in EB])
    return net, EW, EB


@pytest.mark.parametrize("name", NETS)
def test_parallel_mlp_reproduces_reference_vectors(name):
    c = CASES[name]
    net, EW, EB = _net_from_case(c)
    pw, pb = net.perturbations()            # what is stored == what the reference was given (fp16-representable)
    for i in range(len(EW)):
        assert np.array_equal(pw[i].cpu().numpy(), EW[i]) and np.array_equal(pb[i].cpu().numpy(), EB[i])
    actions = net.forward(torch.from_numpy(c["obs"]).cuda())
    np.testing.assert_allclose(actions.cpu().numpy(), c["actions"], rtol=RTOL, atol=ATOL)
    for u in range(2):
        if u:
            net.set_perturbations([torch.from_numpy(e) for e in EW], [torch.from_numpy(e) for e in EB])
        net.reconstruct_perturbations()
        net.update_parameters(torch.from_numpy(c[f"fitness{u}"]).cuda())
        for i in range(len(EW)):
            np.testing.assert_allclose(net.weight_layers[i].cpu().numpy(), c[f"w{i}_after{u}"], rtol=2e-5, atol=1e-6)
            np.testing.assert_allclose(net.bias_layers[i].cpu().numpy(), c[f"b{i}_after{u}"], rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(net.get_l2_norm(), float(c["l2_norm"]), rtol=1e-5)
    with pytest.raises(RuntimeError):       # like the reference, forward needs fresh perturbations after an update
        net.forward(torch.from_numpy(c["obs"]).cuda())


@pytest.mark.parametrize("shape,N,E,atol", [((60, 8, 1), 20002, 2, ATOL), ((60, 16, 4, 2), 3000, 0, ATOL), ((60, 1), 4100, 100, ATOL),
                                            ((60, 300, 2), 520, 8, 1e-5),    # 300-term f32 dot products of O(1) terms
                                            ((400, 12, 2), 600, 4, 1e-5),    # > 319 inputs: the generic kernel
                                            ((400, 64, 2), 300, 2, 1e-5),    # ... with theta too large for shared memory
                                            ((300, 64, 64, 1), 2050, 2, 1e-5)])   # streaming kernel, several chunks per layer
def test_forward_vs_oracle_large_and_lazy_equals_dense(shape, N, E, atol):
    """Many pairs per warp (persistent loop), eval envs, several chunks per layer, all three forward kernels (fast:
    5R -> <= 8 first layer; streaming: layers of <= 319 inputs; generic: anything); forward from lazy handles ==
    forward from the materialised tensor, bit for bit."""
    from finenvs_b200.agents.networks import ParallelMLP
    from oracle import oracle_es as oes

    W = shape[0] // 5
    series = _series(W)
    env = _env(series, N, flat_obs=True)
    torch.manual_seed(11)
    net = ParallelMLP(N, E, shape, noise_std_dev=0.3, device_id=0, seed=9, action_noise_std=0.0)
    net.perturb_parameters()
    lo = env.reset_lazy()
    for t in range(3):
        dense = lo.materialize()
        a_lazy, a_dense = net.forward(lo), net.forward(dense)
        assert torch.equal(a_lazy, a_dense)
        pw, pb = net.perturbations()
        ref = oes.forward([w.cpu().numpy() for w in net.weight_layers], [b.cpu().numpy() for b in net.bias_layers],
                          [e.cpu().numpy() for e in pw], [e.cpu().numpy() for e in pb], 0.3, E, dense.cpu().numpy())
        np.testing.assert_allclose(a_lazy.cpu().numpy(), ref, rtol=RTOL, atol=atol)
        lo, _, _, _ = env.step_lazy(a_lazy[:, :1].contiguous())


def test_perturbations_are_keyed_and_standard_normal():
    from finenvs_b200.agents.networks import ParallelMLP

    shape, N = (300, 8, 1), 4096
    torch.manual_seed(0)
    a = ParallelMLP(N, 0, shape, device_id=0, seed=42)
    a.perturb_parameters()
    e1 = a._eps.clone()
    pw, pb = a.perturbations()
    x = torch.cat([pw[0].reshape(-1), pb[0].reshape(-1), pw[1].reshape(-1), pb[1].reshape(-1)]).double()
    assert abs(float(x.mean())) < 2e-3 and abs(float(x.var()) - 1.0) < 5e-3
    assert abs(float((x ** 4).mean()) - 3.0) < 0.05 and float(x.abs().max()) < 6.0
    a.perturb_parameters()
    assert not torch.equal(a._eps, e1)                       # a new generation
    torch.manual_seed(0)
    b = ParallelMLP(N, 0, shape, device_id=0, seed=42)
    b.perturb_parameters()
    assert torch.equal(b._eps, e1)                           # same key, same generation -> same perturbations
    # a shard holding global pairs [600, 600+500) regenerates exactly that slice
    torch.manual_seed(0)
    c = ParallelMLP(1000, 0, shape, device_id=0, seed=42, pair_id_base=600, total_pairs=N // 2)
    c.perturb_parameters()
    assert torch.equal(c._eps, e1[600:1100])
    torch.manual_seed(0)
    d = ParallelMLP(N, 0, shape, device_id=0, seed=43)
    d.perturb_parameters()
    assert not torch.equal(d._eps, e1)


def test_action_noise_semantics():
    """parallel_mlp.py:104-109: N(0, 0.01) on the training envs, none on the eval envs — and none at all when there
    are no eval envs (the reference's `action_noise[-0:, :] = 0`)."""
    from finenvs_b200.agents.networks import ParallelMLP

    shape, N = (20, 4, 2), 40000
    obs = torch.randn((N, 20), device="cuda")
    outs = {}
    for E, std in [(0, 0.0), (0, 0.01), (100, 0.0), (100, 0.01)]:
        torch.manual_seed(1)
        net = ParallelMLP(N, E, shape, device_id=0, seed=3, action_noise_std=std)
        net.perturb_parameters()
        outs[(E, std)] = (net.forward(obs), net.forward(obs))
    assert torch.equal(outs[(0, 0.0)][0], outs[(0, 0.01)][0])
    clean, noisy, noisy2 = outs[(100, 0.0)][0], outs[(100, 0.01)][0], outs[(100, 0.01)][1]
    assert torch.equal(clean[-100:], noisy[-100:])
    diff = (noisy - clean)[:-100].double()
    assert abs(float(diff.mean())) < 3e-4 and abs(float(diff.std()) - 0.01) < 2e-4
    assert abs(float(torch.corrcoef(diff.T)[0, 1])) < 0.02                      # the two actions get independent noise
    assert not torch.equal(noisy, noisy2)                                        # fresh noise every call


@pytest.mark.parametrize("name", ACCTS)
def test_evo_agent_accounting_reproduces_reference(name):
    from finenvs_b200.agents.ES import EvoAgent

    c = CASES[name]
    T, N = c["rewards"].shape
    E = int(c["num_eval"])
    if (N - E) % 2:
        pytest.skip("odd training population")
    torch.manual_seed(7)
    agent = EvoAgent({"env_name": "x", "num_envs": N, "num_eval_envs": E, "num_observations": 6, "num_actions": 1},
                     hidden_dims=(4,), write_to_csv=False, device_id=0)
    for t in range(T):
        n, ts = agent.store(torch.from_numpy(c["rewards"][t]).cuda(), torch.from_numpy(c["dones"][t]).cuda())
        assert (int(n), int(ts)) == tuple(int(x) for x in c["counts"][t])
        assert (n >= 1) == (c["counts"][t][0] >= 1)
    assert np.array_equal(agent.finished_returns.cpu().numpy(), c["finished_returns"])
    assert np.array_equal(agent.dones.cpu().numpy(), c["done_envs"])
    assert np.array_equal(agent.current_returns.cpu().numpy(), c["current_returns"])
    agent.perform_rank_transformation()
    agent.compute_mean_returns()
    # torch on CUDA divides by a scalar as a multiplication by its reciprocal: last-bit differences from the CPU run
    np.testing.assert_allclose(agent.centered_ranks.cpu().numpy(), c["centered_ranks"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(agent.final_ranks.cpu().numpy(), c["final_ranks"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(agent.mean_returns.cpu().numpy(), c["mean_returns"], rtol=1e-6, atol=1e-7)


def test_es_training_loop_as_in_the_reference_example():
    """examples/isaac_gym/ES_MLP_Isaac_Gym.py:30-38 verbatim on the trading env with lazy observations; the mean
    training return must improve over generations on a series with an exploitable drift."""
    from finenvs_b200.agents.ES import EvoAgent
    from finenvs_b200.data import loader
    from finenvs_b200.environments import TimeSeriesEnv

    W, bars, days = 4, 30, 60
    rng = np.random.default_rng(0)
    n = bars * days
    close = 100.0 * np.exp(np.cumsum(np.full(n, 0.002)))      # steady up-trend: being long pays
    opn = np.concatenate([[100.0], close[:-1]])
    prices = np.round(np.stack([opn, np.maximum(opn, close) * 1.0005, np.minimum(opn, close) * 0.9995, close], 1), 4)
    seg_start, seg_len = loader.regular_segments(n, bars, W)
    series = loader.stage_series(prices, seg_start, seg_len, W, "cuda:0", torch.float32)
    num_envs, num_eval_envs, episodes_per_batch = 4096 + 12, 12, 8000
    env = TimeSeriesEnv("trend", num_intervals=W, device_id=0, series=series, num_envs=num_envs, seed=1,
                        random_reset="all", flat_obs=True, num_eval_envs=num_eval_envs)
    torch.manual_seed(0)
    agent = EvoAgent(env.get_env_args(), hidden_dims=(8,), learning_rate=0.05, noise_std_dev=0.1, write_to_csv=False, seed=4)
    states = env.reset_all(lazy=True)
    history = []
    steps = 0
    while len(history) < 6 and steps < 4000:
        actions = agent.step(states)
        (next_states, rewards, dones, _) = env.step_lazy(actions)
        (num_done, _) = agent.store(rewards, dones)
        states = next_states
        steps += 1
        if num_done >= episodes_per_batch:
            agent.compute_mean_returns()
            history.append(float(agent.mean_returns[: agent.num_training_envs].mean()))
            agent.train()
            states = env.reset_all(lazy=True)
    assert len(history) == 6, steps
    assert history[-1] > history[0] + 1.0, history
