"""Pins the CPU oracle (oracle/fe_oracle.c) to traces the real reference produced
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import oracle as orc
from parity_utils import golden_traces, load_trace, oracle_state, replay_trace, trace_params, trace_series


@pytest.mark.parametrize("name", golden_traces())
@pytest.mark.parametrize("out_f64,multi", [(True, False), (False, False), (True, True)])
def test_oracle_replays_reference_trace(name, out_f64, multi):
    """multi=True runs the multi-asset extension's code path at A = 1: it must still be the reference."""
    z = load_trace(name)
    fs = trace_series(z)
    env = orc.OracleEnv(fs, num_envs=len(z["seg_init"]), evaluate=bool(z["evaluate"]), seed=int(z["seed"]),
                        seg_init=z["seg_init"], out_f64=out_f64, force_multi=multi, **trace_params(z))
    replay_trace(z, env, lambda: oracle_state(env), out_f64, name)
    # the redraws the reference consumed are the oracle's own Philox draws
    for step, kind, seg in z["draw_log"]:
        env_id = fs.num_segments if kind == 1 else len(z["seg_init"]) - 1   # ctor draw: env id D (:253)
        r = orc.philox(int(z["seed"]), env_id, int(step), int(kind))
        assert (int(r[0]) * fs.num_segments) >> 32 == seg


def test_kat_matches_survey_appendix_b():
    """IBM dummy, W=390, five scripted steps."""
    z = load_trace("kat_ibm_w390.npz")
    assert z["obs_reset"].shape == (3, 390, 5)
    assert abs(z["obs_reset"].sum() - 0.9396823296) < 1e-9
    np.testing.assert_allclose(z["rewards"][0], [-0.04999999702, -0.04999999702, -0.01999999955], rtol=1e-9)
    np.testing.assert_array_equal(z["states"][4][2].astype(np.float32),
                                  np.array([9791.0390625, 9614.0390625, 9910.208984375], np.float32))
    np.testing.assert_allclose(z["states"][4][5], [0, 0, 89.55], rtol=1e-12)


def test_adversarial_traces_reach_the_rare_branches():
    """Margin calls, bankruptcies and ragged-day dones must actually occur in the golden set."""
    z = load_trace("trace_adv_s15_w8_n96.npz")
    st = z["states"]
    ptr_reset = (st[:, 1] == 0)
    assert z["dones"].sum() > 1000
    # bankruptcy: done while the pointer was nowhere near the end of the shortest segment
    fs = trace_series(z)
    assert fs.seg_len_raw.min() == 9 and fs.seg_len_raw.max() == 48
    # rewards below -50 only happen through margin calls (a 5-share P&L step is far smaller)
    assert (z["rewards"] < -50).any()
    assert ptr_reset.any()
