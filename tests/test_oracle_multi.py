"""Config 5 (several senders on one bottleneck): the oracle's multi-sender path against the reference's own
Network / Link / Sender classes, driven with 2-3 senders and the external patch Sender.__lt__ = id order
(without it the reference raises TypeError on cross-sender exact ties).  Build container only."""
import numpy as np
import pytest

import oracle
import refharness as rh
from philox_py import PhiloxStream

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="no /root/reference here")


class RefMulti(object):
    """The env glue of network_sim.py:406-484 written out for S senders around the reference classes."""

    def __init__(self, ns, stream, bw, lat, queue, loss, rates, features):
        self.ns = ns
        ns.Sender.__lt__ = lambda a, b: a.id < b.id          # external patch, no reference file is modified
        ns.random = rh.StreamShim([stream])
        links = [ns.Link(bw, lat, queue, loss), ns.Link(bw, lat, queue, loss)]
        self.senders = [ns.Sender(r, [links[0], links[1]], 0, features, history_len=10) for r in rates]
        self.run_dur = 3 * lat
        self.net = ns.Network(self.senders, links)
        self.net.run_for_dur(self.run_dur)
        self.net.run_for_dur(self.run_dur)

    def step(self, actions):
        for s, a in zip(self.senders, actions):
            s.apply_rate_delta(a)
        r0 = self.net.run_for_dur(self.run_dur)
        out = []
        for i, s in enumerate(self.senders):
            s.record_run()
            obs = np.array(s.get_obs()).reshape(-1)
            mi = s.get_run_data()
            rew = (10.0 * mi.get("recv rate") / (8 * 1500) - 1e3 * mi.get("avg latency") - 2e3 * mi.get("loss ratio")) * 0.001
            if i == 0:
                assert rew == r0
                avg0 = mi.get("avg latency")
            mi.get("latency ratio")
            out.append((obs, rew, (s.sent, s.acked, s.lost)))
        if avg0 > 0.0:
            self.run_dur = 0.5 * avg0
        return out


@pytest.mark.parametrize("seed,n_senders", [(1, 2), (2, 2), (3, 3), (4, 2)])
def test_oracle_multi_sender_matches_patched_reference(seed, n_senders):
    ns = rh.load_reference()
    g = np.random.default_rng(seed)
    real_random, had_lt = ns.random, getattr(ns.Sender, "__lt__", None)
    feats = oracle.DEFAULT_FEATURES.split(",")
    try:
        for trial in range(3):
            # config 5 grid: bw 1..1000 Mbit/s (83..83 333 packets/s), delay 1..500 ms; also overdriven links
            bw = float(np.exp(g.uniform(np.log(83.0), np.log(83333.0)))) if trial else 150.0
            lat = float(np.exp(g.uniform(np.log(0.001), np.log(0.5))))
            queue = 1 + int(np.exp(g.uniform(0, 6)))
            loss = float(g.choice([0.0, 0.01, 0.05]))
            rates = [float(g.uniform(40, 900)) for _ in range(n_senders)]
            ref = RefMulti(ns, PhiloxStream(100 + seed), bw, lat, queue, loss, rates, feats)
            o = oracle.OracleEnv()
            o.seed_philox(100 + seed)
            o.reset_multi(bw, lat, queue, loss, rates)
            assert o.cur_time == ref.net.cur_time
            for t in range(120):
                acts = g.normal(0, 2.0, n_senders)
                want = ref.step([float(a) for a in acts])
                obs, rew, done, cnt = o.step_multi(acts)
                for i in range(n_senders):
                    assert tuple(cnt[i]) == want[i][2], (trial, t, i)
                    assert rew[i] == want[i][1] and np.array_equal(obs[i], want[i][0]), (trial, t, i)
                assert o.cur_time == ref.net.cur_time and o.run_dur == ref.run_dur
    finally:
        ns.random = real_random
        if had_lt is None:
            try:
                del ns.Sender.__lt__
            except AttributeError:
                pass


@pytest.mark.parametrize("n_senders", [1, 3, 4])
def test_oracle_multi_ragged_corners_match_patched_reference(n_senders):
    """The corners the round-2 GPU tests lean on the oracle for (tests/test_gpu_multi.py, 1 / 3 / 4 senders on ragged
    links): queues of 0, 1 and 2 packets, loss 0 / 0.05 / 1, very short and long delays -- the oracle against the live
    reference classes."""
    ns = rh.load_reference()
    g = np.random.default_rng(50 + n_senders)
    real_random, had_lt = ns.random, getattr(ns.Sender, "__lt__", None)
    feats = oracle.DEFAULT_FEATURES.split(",")
    try:
        for trial in range(8):
            bw = float(g.uniform(80, 2000))
            lat = float(np.exp(g.uniform(np.log(0.002), np.log(0.3))))
            queue = int(g.integers(0, 3)) if trial % 2 == 0 else int(g.integers(0, 60))
            loss = float(g.choice([0.0, 0.05, 1.0]))
            rates = [float(g.uniform(40, 1500)) for _ in range(n_senders)]
            ref = RefMulti(ns, PhiloxStream(300 + trial), bw, lat, queue, loss, rates, feats)
            o = oracle.OracleEnv()
            o.seed_philox(300 + trial)
            o.reset_multi(bw, lat, queue, loss, rates)
            assert o.cur_time == ref.net.cur_time
            for t in range(40):
                acts = g.normal(0, 2.0, n_senders)
                want = ref.step([float(a) for a in acts])
                obs, rew, done, cnt = o.step_multi(acts)
                for i in range(n_senders):
                    assert tuple(cnt[i]) == want[i][2], (trial, t, i, queue, loss)
                    assert rew[i] == want[i][1] and np.array_equal(obs[i], want[i][0]), (trial, t, i, queue, loss)
                assert o.cur_time == ref.net.cur_time and o.run_dur == ref.run_dur
    finally:
        ns.random = real_random
        if had_lt is None:
            try:
                del ns.Sender.__lt__
            except AttributeError:
                pass
