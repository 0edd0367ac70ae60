"""The CPU oracle against the unmodified reference imported from /root/reference -- runs only in
the build container (skipped on the GPU box, where the tree does not exist)."""
import random

import numpy as np
import pytest

import oracle
import refharness as rh

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="no /root/reference here")


@pytest.mark.parametrize("seed", [3, 77, 31337])
def test_live_reference_mt(seed):
    ns = rh.load_reference()
    arng = random.Random(seed + 1)
    e = oracle.OracleEnv()
    e.seed_mt(seed)
    e.sample_params()
    with rh.quiet_tmp_cwd():
        random.seed(seed)
        env = ns.SimulatedNetworkEnv()
        for ep in range(3):
            obs0 = env.reset()
            p = e.sample_params()
            l, s = env.links[0], env.senders[0]
            assert (l.bw, l.dl, l.lr, s.starting_rate) == (p[0], p[1], p[3], p[4])
            assert np.array_equal(e.reset(*p), obs0)
            for t in range(400):
                a = arng.gauss(0, 1.5)
                obs, r, d, _ = env.step([a])
                o2, r2, d2, c2, _ = e.step(a)
                s = env.senders[0]
                assert (s.sent, s.acked, s.lost) == tuple(c2), (ep, t)
                assert np.array_equal(obs, o2) and float(r) == r2 and d == d2, (ep, t)
                assert env.net.cur_time == e.cur_time and float(env.run_dur) == e.run_dur
                assert len(env.net.q) == e.L.pcco_queue_len(e.h)


def test_live_reference_lockstep_philox_envs():
    """Several reference envs in lock step, each on its own Philox stream (StreamShim)."""
    from philox_py import PhiloxStream
    ns = rh.load_reference()
    n = 6
    streams = [PhiloxStream(1000 + i) for i in range(n)]
    shim = rh.StreamShim(streams)
    g = np.random.default_rng(9)
    real = ns.random
    ns.random = shim
    try:
        with rh.quiet_tmp_cwd():
            envs, orcs = [], []
            for i in range(n):
                shim.select(i)
                shim.script = [100.0, 0.1, 0.0, 0.0, 1.0]
                envs.append(ns.SimulatedNetworkEnv())
                o = oracle.OracleEnv()
                o.seed_philox(1000 + i)
                orcs.append(o)
            for i in range(n):
                shim.select(i)
                obs0 = envs[i].reset()          # parameters drawn from stream i by the reference
                p = orcs[i].sample_params()
                l, s = envs[i].links[0], envs[i].senders[0]
                assert (l.bw, l.dl, l.lr, s.starting_rate) == (p[0], p[1], p[3], p[4])
                assert np.array_equal(orcs[i].reset(*p), obs0)
            for t in range(150):
                acts = g.normal(0, 2, size=n)
                for i in range(n):
                    shim.select(i)
                    obs, r, d, _ = envs[i].step([float(acts[i])])
                    o2, r2, d2, c2, _ = orcs[i].step(float(acts[i]))
                    assert np.array_equal(obs, o2) and float(r) == r2, (i, t)
                    s = envs[i].senders[0]
                    assert (s.sent, s.acked, s.lost) == tuple(c2)
    finally:
        ns.random = real


def test_live_reference_one_and_two_packet_queues():
    """Scripted link parameters with queue = 1 and 2 packets (1 + int(exp(x)) with exp(x) < 1 resp. < 2), the corner
    of pcc_core.cuh's tail-drop threshold: the oracle against the live reference env, 12 links x 80 steps."""
    import math
    from philox_py import PhiloxStream
    ns = rh.load_reference()
    g = np.random.default_rng(21)
    real = ns.random
    try:
        with rh.quiet_tmp_cwd():
            for i in range(12):
                shim = rh.StreamShim([PhiloxStream(2000 + i)])
                ns.random = shim
                shim.script = [100.0, 0.1, 0.0, 0.0, 1.0]
                env = ns.SimulatedNetworkEnv()
                o = oracle.OracleEnv()
                o.seed_philox(2000 + i)
                queue = 1 + (i % 2)
                bw = float(np.exp(g.uniform(np.log(40), np.log(5000))))
                lat = float(np.exp(g.uniform(np.log(0.002), np.log(0.5))))
                loss, factor = float(g.choice([0.0, 0.02, 0.3])), float(g.uniform(0.3, 3.0))
                shim.script = [bw, lat, math.log(queue - 1 + 0.5), loss, factor]
                obs0 = env.reset()
                l, s = env.links[0], env.senders[0]
                assert int(round(l.max_queue_delay * l.bw)) == queue
                assert np.array_equal(o.reset(bw, lat, queue, loss, s.starting_rate), obs0)
                for t in range(80):
                    a = float(g.normal(0, 3.0))
                    obs, r, d, _ = env.step([a])
                    o2, r2, d2, c2, _ = o.step(a)
                    assert (s.sent, s.acked, s.lost) == tuple(c2), (i, t)
                    assert np.array_equal(obs, o2) and float(r) == r2, (i, t)
                    assert env.net.cur_time == o.cur_time and float(env.run_dur) == o.run_dur
    finally:
        ns.random = real
