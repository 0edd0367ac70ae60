"""The product's scalar ingestion core (pcc-rl_b200/csrc/pcc_flows_core.cuh) compiled for the host,
against the reference's committed outputs (tests/golden/flows_*.npz).  Test harness, not a fallback."""
import numpy as np
import pytest

import twin_util
from flows_util import CLIENT_RATE, SHIM_RATE, load_flows_golden, record_of


@pytest.mark.parametrize("name,rate_cfg", [("flows_client_default", CLIENT_RATE), ("flows_allfeatures", None),
                                           ("flows_shim", SHIM_RATE)])
def test_twin_flows_golden(name, rate_cfg):
    g = load_flows_golden(name)
    flows = [twin_util.TwinFlow(g["history_len"], g["features"]) for _ in range(g["n_flows"])]
    rates = [rate_cfg["start"]] * g["n_flows"] if rate_cfg else None
    for k in range(len(g["op"])):
        i = int(g["flow"][k])
        if g["op"][k] == 1:
            flows[i].reset(g["reset_mode"])
            if rate_cfg and rate_cfg["style"] == 1:
                rates[i] = rate_cfg["start"]
        else:
            if rate_cfg and rate_cfg["style"] == 1:
                rates[i] = flows[i].apply_rate_delta(rates[i], g["action"][k], rate_cfg)
            m = flows[i].give_sample(record_of(g, k))
            if not np.isnan(g["metrics"][k]).any():
                assert np.array_equal(m, g["metrics"][k]), "metrics of op %d" % k
            if rate_cfg and rate_cfg["style"] == 0:
                rates[i] = flows[i].apply_rate_delta(rates[i], g["action"][k], rate_cfg)
                assert rates[i] * 1e6 == g["rate"][k]
        assert np.array_equal(flows[i].obs(), g["obs"][k]), "obs after op %d" % k
        if rate_cfg and (rate_cfg["style"] == 1 or g["op"][k] == 1):
            assert rates[i] == g["rate"][k]


def test_lane_level_summation_spec_equals_numpy():
    """The 8-lane / xor-tree / flat-or-stack summation the flows kernels implement (emulated lane by lane in
    tests/twin/pcc_twin.cpp) reproduces numpy's pairwise sums bit for bit for every n up to the subgroup limit."""
    import ctypes as C
    L = twin_util.lib()
    L.twin_flows_pass_sums.argtypes = [C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double)]
    g = np.random.default_rng(0)
    a = g.uniform(0.01, 0.7, 4096) * np.exp(g.normal(0, 2, 4096))       # wide dynamic range: order matters
    out = np.zeros(3)
    for n in list(range(1, 600)) + [1023, 1024, 1025, 1500, 1799, 1800]:
        off = int(g.integers(0, 4096 - n))
        x = np.ascontiguousarray(a[off:off + n])
        L.twin_flows_pass_sums(x.ctypes.data_as(C.POINTER(C.c_double)), n, out.ctypes.data_as(C.POINTER(C.c_double)))
        half = n // 2
        assert (0.0 + out[0]) / n == np.mean(x), n
        if half >= 1:
            assert (0.0 + out[1]) / half == np.mean(x[:half]) and (0.0 + out[2]) / (n - half) == np.mean(x[half:]), n
