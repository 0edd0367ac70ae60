"""Host-side link parameters: the sampler's formulas (gym/network_sim.py:454-467) stay inside what the device loops
assume, and explicit parameters that would stall a pacing timer are refused before they reach the GPU."""
import numpy as np
import pytest

from pcc_rl_b200.params import LinkRanges, sample_link_params, validate_link_params


def test_sampled_params_are_valid_and_shard_invariant():
    p = sample_link_params(seed=3, episode=2, global_ids=np.arange(1000), n_global=1000)
    validate_link_params(p["bw"], p["lat"], p["queue"], p["loss"], p["start_rate"])
    r = LinkRanges()
    assert p["queue"].min() >= 2 and p["queue"].max() <= r.max_queue_packets()       # 1 + int(exp(U(0, 8)))
    assert (p["bw"] >= r.bw[0]).all() and (p["bw"] <= r.bw[1]).all()
    half = sample_link_params(seed=3, episode=2, global_ids=np.arange(500, 1000), n_global=1000)
    for k in p:
        assert np.array_equal(p[k][500:], half[k])


@pytest.mark.parametrize("field,value", [("bw", 0.0), ("bw", np.nan), ("bw", -5.0), ("lat", -1e-3), ("lat", np.inf),
                                         ("queue", -1), ("loss", 1.5), ("loss", -0.1), ("start_rate", 0.0),
                                         ("start_rate", -40.0), ("start_rate", np.nan)])
def test_out_of_range_params_are_refused(field, value):
    p = dict(bw=np.full(4, 200.0), lat=np.full(4, 0.1), queue=np.full(4, 10), loss=np.full(4, 0.01),
             start_rate=np.full(4, 150.0))
    validate_link_params(p["bw"], p["lat"], p["queue"], p["loss"], p["start_rate"])
    p[field] = p[field].astype(np.float64 if field != "queue" else np.int64)
    p[field][2] = value
    with pytest.raises(ValueError, match=field):
        validate_link_params(p["bw"], p["lat"], p["queue"], p["loss"], p["start_rate"])


def test_corner_values_are_accepted():
    """queue 0 / 1, loss 0 / 1, zero delay: legal for the reference's classes, covered by the tiny-queue parity tests."""
    validate_link_params([83.3, 1e5], [0.0, 0.5], [0, 1], [0.0, 1.0], [[40.0, 1000.0], [1.0, 2.0]])
