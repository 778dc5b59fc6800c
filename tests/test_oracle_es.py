"""The ES-path oracle (oracle/oracle_es.py) against vectors produced by the reference's own ParallelMLP / EvoAgent
(tests/golden/es_path.npz, generator tests/golden/make_golden_es.py)."""
import os

import numpy as np
import pytest

from oracle import oracle_es as oes

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "es_path.npz")


def es_cases():
    with np.load(GOLD) as z:
        names = sorted({k.split(".")[0] for k in z.files})
        return {n: {k.split(".", 1)[1]: z[k] for k in z.files if k.startswith(n + ".")} for n in names}


CASES = es_cases()
NETS = sorted(n for n in CASES if not n.startswith("acct"))
ACCTS = sorted(n for n in CASES if n.startswith("acct"))


def net_arrays(c):
    L = len(c["shape"]) - 1
    return ([c[f"w{i}"].copy() for i in range(L)], [c[f"b{i}"].copy() for i in range(L)],
            [c[f"eps_w{i}"] for i in range(L)], [c[f"eps_b{i}"] for i in range(L)])


@pytest.mark.parametrize("name", NETS)
def test_forward_and_update_match_reference(name):
    c = CASES[name]
    W, B, EW, EB = net_arrays(c)
    E, sigma = int(c["num_eval"]), float(c["sigma"])
    a = oes.forward(W, B, EW, EB, sigma, E, c["obs"])
    np.testing.assert_allclose(a, c["actions"], rtol=1e-5, atol=1e-6)   # f32 dot products: summation order differs
    adam = oes.Adam(W, B, 0.01)
    for u in range(2):
        oes.update_parameters(W, B, EW, EB, adam, c[f"fitness{u}"], E, 0.005)
        for i in range(len(W)):
            np.testing.assert_allclose(W[i], c[f"w{i}_after{u}"], rtol=2e-5, atol=1e-6)   # Adam steps are ~lr = 1e-2; 1e-6 absolute = 1e-4 of a step
            np.testing.assert_allclose(B[i], c[f"b{i}_after{u}"], rtol=2e-5, atol=1e-6)   # Adam steps are ~lr = 1e-2; 1e-6 absolute = 1e-4 of a step


@pytest.mark.parametrize("name", ACCTS)
def test_accounting_and_ranks_match_reference(name):
    c = CASES[name]
    T, N = c["rewards"].shape
    acc = oes.Accounting(N)
    for t in range(T):
        n, ts = acc.step_and_store(c["rewards"][t], c["dones"][t])
        assert (n, ts) == tuple(c["counts"][t])
    assert np.array_equal(acc.fin_ret, c["finished_returns"]) and np.array_equal(acc.dones, c["done_envs"])
    assert np.array_equal(acc.cur_ret, c["current_returns"])
    centred, final = acc.final_ranks()
    assert len(np.unique(acc.fin_ret)) == len(acc.fin_ret)              # no ties: the rank transform is well defined
    assert np.array_equal(centred, c["centered_ranks"])
    np.testing.assert_allclose(final, c["final_ranks"], rtol=0, atol=1e-6)   # index_add_ accumulation order
    np.testing.assert_allclose(acc.mean_returns(), c["mean_returns"], rtol=1e-6, atol=1e-7)
