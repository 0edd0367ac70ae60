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
