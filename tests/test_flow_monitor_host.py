"""Host-side pieces of the ingestion path that need no GPU: the shim's text wire format
(udt-plugins/training/shim.py:31-42 writes it, gym/online/shim_env.py:108-121 reads it), the feature tables,
and the loud failure without a CUDA device."""
import numpy as np
import pytest

import pcc_rl_b200
from pcc_rl_b200 import flow_monitor as fm
from flows_util import load_flows_golden, record_of


def test_wire_format_round_trip_on_the_reference_stream():
    """Every record the reference shim sent over its socket (tests/golden/flows_shim.npz): format -> parse gives the
    fields the reference's ShimNetworkEnv.step parsed (they were quantised by the same "%f")."""
    g = load_flows_golden("flows_shim")
    for k in np.nonzero(g["op"] == 0)[0]:
        r = record_of(g, int(k))
        line = fm.format_sample_line(31, r["bytes_sent"], r["bytes_acked"], r["bytes_lost"], r["send_start"], r["send_end"],
                                     r["recv_start"], r["recv_end"], [float(v) for v in r["rtt"]], r["packet_size"], 1.5)
        assert line.endswith("\n") and line.count(";") == 10
        p = fm.parse_sample_line("stale;line\n" + line)            # like the reference: the LAST complete line wins
        assert p["flow_id"] == 31 and p["utility"] == 1.5
        assert (p["bytes_sent"], p["bytes_acked"], p["bytes_lost"], p["packet_size"]) == (
            r["bytes_sent"], r["bytes_acked"], r["bytes_lost"], r["packet_size"])
        assert (p["send_start_time"], p["send_end_time"], p["recv_start_time"], p["recv_end_time"]) == (
            r["send_start"], r["send_end"], r["recv_start"], r["recv_end"])
        assert p["rtt_samples"] == [float(v) for v in r["rtt"]]


def test_wire_format_matches_shim_py_literally():
    line = fm.format_sample_line(7, 3000, 1500, 0, 0.25, 0.5, 0.3, 0.55, [0.05, 0.051], 1500, -2.0)
    assert line == "7;3000;1500;0;0.250000;0.500000;0.300000;0.550000;[0.05, 0.051];1500;-2.000000\n"


def test_rate_constants_of_both_callers():
    assert fm.CLIENT_DEFAULTS == dict(delta_scale=0.05, min_rate=0.5, max_rate=300.0, rate_style=0)   # loaded_client.py:33-35
    assert fm.SHIM_DEFAULTS == dict(delta_scale=0.025, min_rate=0.25, max_rate=1000.0, rate_style=1)  # shim_env.py:38-44
    assert fm.RESET_RATE_MIN == fm.RESET_RATE_MAX == 6.0 and fm.STARTING_RATE == 2.0


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError):
        pcc_rl_b200.PccFlowMonitor(16)
    with pytest.raises(RuntimeError):
        pcc_rl_b200.PccMultiSenderEnv(4, n_senders=1, use_cwnd=True)


def test_flows_config_defaults_through_the_abi():
    """pcc_flows_default_config / pcc_flows_workspace_bytes / pcc_default_variant need no device."""
    import ctypes as C
    from pcc_rl_b200 import _lib
    L = _lib.load()
    cfg = _lib.PccFlowsConfig()
    L.pcc_flows_default_config(C.byref(cfg))
    assert (cfg.history_len, cfg.n_features, list(cfg.feature_ids)[:3]) == (10, 3, [7, 10, 11])
    assert (cfg.delta_scale, cfg.min_rate, cfg.max_rate, cfg.rate_style) == (0.05, 0.5, 300.0, 0)
    cfg.n_flows = 1000
    nb = C.c_uint64()
    assert L.pcc_flows_workspace_bytes(C.byref(cfg), C.byref(nb)) == 0
    assert nb.value >= 1000 * (30 * 8 + 32)
    cfg.history_len = 0
    assert L.pcc_flows_workspace_bytes(C.byref(cfg), C.byref(nb)) == _lib.PCC_EINVAL
    v = _lib.PccVariant()
    L.pcc_default_variant(C.byref(v))
    assert (v.use_cwnd, v.use_latency_noise, v.max_latency_noise, v.initial_cwnd, v.min_cwnd, v.max_cwnd) == (0, 0, 1.1, 25, 4, 5000)
