"""MI-sample ingestion path: the CPU oracle (oracle/pcc_oracle_flows.c) against
outputs of the unmodified reference modules -- committed (tests/golden/flows_*.npz) and live."""
import numpy as np
import pytest

import oracle
import refharness
from flows_util import CLIENT_RATE, SHIM_RATE, load_flows_golden, record_of

NAMES = ["flows_client_default", "flows_allfeatures", "flows_shim"]


def replay_oracle(g, rate_cfg=None):
    fl = oracle.OracleFlows(g["n_flows"], g["history_len"], g["features"])
    rates = [rate_cfg["start"]] * g["n_flows"] if rate_cfg else None
    for k in range(len(g["op"])):
        i = int(g["flow"][k])
        if g["op"][k] == 1:
            fl.reset(i, g["reset_mode"])
            if rate_cfg and rate_cfg["style"] == 1:
                rates[i] = rate_cfg["start"]            # ShimNetworkEnv.reset: set_rate(STARTING_RATE)
        else:
            r = record_of(g, k)
            if rate_cfg and rate_cfg["style"] == 1:     # shim: the action is applied BEFORE the MI is read
                rates[i] = fl.apply_rate_delta(rates[i], g["action"][k], rate_cfg["delta_scale"],
                                               rate_cfg["min_rate"], rate_cfg["max_rate"], 1)
            m = fl.give_sample(i, r["bytes_sent"], r["bytes_acked"], r["bytes_lost"], r["send_start"], r["send_end"],
                               r["recv_start"], r["recv_end"], r["rtt"], r["packet_size"], want_metrics=True)
            if not np.isnan(g["metrics"][k]).any():
                assert np.array_equal(m, g["metrics"][k]), "metrics of op %d" % k
            if rate_cfg and rate_cfg["style"] == 0:     # client: get_rate() after the sample
                rates[i] = fl.apply_rate_delta(rates[i], g["action"][k], rate_cfg["delta_scale"],
                                               rate_cfg["min_rate"], rate_cfg["max_rate"], 0)
        assert np.array_equal(fl.obs(i), g["obs"][k]), "obs after op %d" % k
        if rate_cfg:
            want = g["rate"][k] if rate_cfg["style"] == 1 or g["op"][k] == 1 else g["rate"][k] / 1e6
            got = rates[i]
            if rate_cfg["style"] == 0 and g["op"][k] == 0:
                assert got * 1e6 == g["rate"][k], "rate after op %d" % k     # get_rate returns rate * 1e6
            else:
                assert got == want, "rate after op %d" % k


def test_flows_golden_present():
    for n in NAMES:
        g = load_flows_golden(n)
        assert len(g["op"]) > 50 and (g["op"] == 1).any()


def test_oracle_flows_client_golden():
    replay_oracle(load_flows_golden("flows_client_default"), CLIENT_RATE)


def test_oracle_flows_allfeatures_golden():
    g = load_flows_golden("flows_allfeatures")
    assert set(np.diff(g["rtt_off"])[g["op"] == 0]) >= {0, 1, 7, 8, 9, 128, 129, 1024, 1025, 4001}
    replay_oracle(g)


def test_oracle_flows_shim_golden():
    replay_oracle(load_flows_golden("flows_shim"), SHIM_RATE)


@pytest.mark.reference
@pytest.mark.skipif(not refharness.reference_available(), reason="reference tree not present")
def test_oracle_flows_vs_live_reference():
    """Random records through common/sender_obs.py (imported unmodified) and the oracle, 12 features."""
    so, _ = refharness.load_reference_flows()
    feats = oracle.METRIC_NAMES
    g = np.random.default_rng(7)
    n_flows, H = 5, 4
    fl = oracle.OracleFlows(n_flows, H, ",".join(feats))
    hist = [so.SenderHistory(H, feats, 555000 + i) for i in range(n_flows)]
    for step in range(400):
        i = int(g.integers(0, n_flows))
        if g.random() < 0.03:
            hist[i] = so.SenderHistory(H, feats, 555000 + i)
            fl.reset(i, 1)
        else:
            n = int(g.integers(0, 600))
            rtt = list(g.uniform(0.01, 0.5) * (1 + g.random(n)))
            t = float(g.uniform(0, 100))
            f = dict(bytes_sent=int(g.integers(0, 10 ** 7)), bytes_acked=int(g.integers(0, 10 ** 7)),
                     bytes_lost=int(g.integers(0, 10 ** 5)), send_start=t, send_end=t + float(g.uniform(0, 0.3)),
                     recv_start=t + 0.05, recv_end=t + 0.05 + float(g.uniform(0, 0.3)),
                     packet_size=int(g.choice([1500, 1200])))
            hist[i].step(so.SenderMonitorInterval(555000 + i, rtt_samples=rtt, **f))
            fl.give_sample(i, f["bytes_sent"], f["bytes_acked"], f["bytes_lost"], f["send_start"], f["send_end"],
                           f["recv_start"], f["recv_end"], rtt, f["packet_size"])
        assert np.array_equal(hist[i].as_array(), fl.obs(i)), "step %d" % step


def test_flow_sharding_equals_the_unsharded_monitor():
    """Flows shard by contiguous id range with no exchange (SURVEY.md §8e applied to the ingestion path): the oracle
    fed with each rank's slice of two global batches (pcc_rl_b200.distributed.shard_flow_batch) reproduces, flow by
    flow, the oracle fed with the whole batches."""
    from pcc_rl_b200 import distributed as D
    from flows_util import synth_batch
    n_flows, world = 1000, 3
    g = np.random.default_rng(21)
    batches = [synth_batch(g, 2500, n_flows, mean_samples=20, unique=False, t0=float(t)) for t in range(2)]
    whole = oracle.OracleFlows(n_flows)
    for b in batches:
        whole.give_batch(b)
    seen = 0
    for rank in range(world):
        lo, hi = D.shard_range(n_flows, rank, world)
        part = oracle.OracleFlows(hi - lo)
        for b in batches:
            sb, rng_ = D.shard_flow_batch(b, rank, world, n_flows)
            assert rng_ == (lo, hi) and sb["flow"].min() >= 0 and sb["flow"].max() < hi - lo
            seen += len(sb["flow"])
            part.give_batch(sb)
        for i in range(lo, hi):
            assert np.array_equal(part.obs(i - lo), whole.obs(i)), (rank, i)
            assert part.conn_min(i - lo) == whole.conn_min(i)
    assert seen == 2 * 2500
