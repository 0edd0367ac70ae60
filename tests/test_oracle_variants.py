"""the congestion-window and latency-noise variants of the event loop (USE_CWND /
USE_LATENCY_NOISE, network_sim.py:51-54, shipped False).  The oracle against outputs of the unmodified
reference with the module switches turned on (tests/golden/variant_*.npz, oracle/gen_golden_variants.py)."""
import numpy as np
import pytest

import oracle
from golden_util import assert_step_equal, golden_names, load_golden, GOLDEN_DIR
import os

SINGLE = [n for n in golden_names("variant_") if not n.startswith("variant_multi_")]
MULTI = golden_names("variant_multi_")


def test_variant_goldens_present():
    assert len(SINGLE) >= 6 and len(MULTI) >= 2


@pytest.mark.parametrize("name", SINGLE)
def test_oracle_variant_golden(name):
    g = load_golden(name)
    use_cwnd, use_noise = bool(g["use_cwnd"]), bool(g["use_noise"])
    e = oracle.OracleEnv(g["history_len"], g["features"])
    e.set_variant(use_cwnd, use_noise)
    e.seed_philox(g["seed"])
    k = 0
    for ep in range(len(g["ep_params"])):
        bw, lat, q, loss, rate = g["ep_params"][ep]
        obs0 = e.reset(bw, lat, int(q), loss, rate)
        assert np.array_equal(obs0, g["ep_obs0"][ep])
        assert e.cur_time == g["ep_cur_time0"][ep]
        for _ in range(g["steps_per_episode"]):
            if use_cwnd:
                obs, r, d, c, info = e.step_cwnd(g["action"][k], g["cwnd_action"][k])
            else:
                obs, r, d, c, info = e.step(g["action"][k])
            assert_step_equal(g, k, obs, r, d, c, e.cur_time, e.run_dur, e.rate, info, name)
            assert e.cwnd() == g["cwnd"][k], "%s step %d cwnd" % (name, k)
            k += 1
    assert k == len(g["action"])
    if use_cwnd:   # the window really bound in these files: fewer packets than rate x duration
        assert (g["cwnd"] != 25).any()


@pytest.mark.parametrize("name", MULTI)
def test_oracle_variant_multi_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    bw, lat, queue, loss = z["params"]
    e = oracle.OracleEnv(10, oracle.DEFAULT_FEATURES)
    e.set_variant(bool(z["use_cwnd"]), bool(z["use_noise"]))
    e.seed_philox(int(z["seed"]))
    e.reset_multi(bw, lat, int(queue), loss, z["rates"])
    assert e.cur_time == float(z["cur_time0"])
    for t in range(len(z["action"])):
        obs, rew, done, cnt = e.step_multi(z["action"][t], z["cwnd_action"][t])
        assert np.array_equal(cnt, z["counts"][t]), "%s step %d counts" % (name, t)
        assert np.array_equal(obs, z["obs"][t]), "%s step %d obs" % (name, t)
        assert np.array_equal(rew, z["reward"][t]), "%s step %d reward" % (name, t)
        assert e.cur_time == z["cur_time"][t] and e.run_dur == z["run_dur"][t]
        assert [e.cwnd(i) for i in range(len(z["rates"]))] == list(z["cwnd"][t])
