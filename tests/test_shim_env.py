"""ShimNetworkEnv (drop-in for gym/online/shim_env.py): the TCP side on the CPU, the whole env against the oracle's
restatement of SenderHistory / the reference's rate formulas on the GPU."""
import socket
import threading

import numpy as np
import pytest

import oracle
from pcc_rl_b200.flow_monitor import format_sample_line


def _sender(port, records, rates_seen, fragment=0, pair_first=False):
    """Plays the UDT sender's shim (udt-plugins/training/shim.py:31-42): read the rate, answer with one record line."""
    s = None
    for _ in range(300):                          # like the real sender: keep trying until the env listens
        try:
            s = socket.create_connection(("localhost", port), timeout=30)
            break
        except OSError:
            import time
            time.sleep(0.1)
    assert s is not None
    s.settimeout(30)
    try:
        for k, rec in enumerate(records):
            rates_seen.append(s.recv(1024).decode())
            line = format_sample_line(**rec).encode()
            if pair_first and k == 0:            # a stale line in front: the env must use the LAST complete line
                line = format_sample_line(**dict(rec, bytes_sent=1, bytes_acked=1, rtt_samples=[9.0])).encode() + line
            if fragment:
                for i in range(0, len(line), fragment):
                    s.sendall(line[i:i + fragment])
            else:
                s.sendall(line)
    finally:
        s.close()


def _records(g, n, max_samples):
    out, t = [], 0.0
    for k in range(n):
        ns = int(g.integers(0, max_samples))
        dur = float(g.uniform(0.02, 0.3))
        sent = int(g.integers(1, 400)) * 1500
        lost = int(g.integers(0, 3)) * 1500
        out.append(dict(flow_id=0, bytes_sent=sent, bytes_acked=max(sent - lost, 0), bytes_lost=lost,
                        send_start_time=round(t, 6), send_end_time=round(t + dur, 6), recv_start_time=round(t + 0.05, 6),
                        recv_end_time=round(t + 0.05 + dur, 6),
                        rtt_samples=[round(float(x), 6) for x in g.uniform(0.04, 0.2, ns)], packet_size=1500,
                        utility=round(float(g.normal(0, 5)), 6)))
        t += dur
    return out


def test_shim_link_reads_whole_records_and_last_complete_line():
    from pcc_rl_b200.shim_env import ShimLink
    g = np.random.default_rng(3)
    recs = _records(g, 6, 400)                       # records of several KB: more than one recv(1024)
    link = ShimLink(port=0)                          # an ephemeral port instead of 9787
    seen = []
    link.sock.settimeout(30)
    th = threading.Thread(target=_sender, args=(link.port, recs, seen), kwargs=dict(fragment=700, pair_first=True), daemon=True)
    th.start()
    try:
        for k, rec in enumerate(recs):
            got = link.exchange(2.0 + k)
            assert got["bytes_sent"] == rec["bytes_sent"] and got["bytes_lost"] == rec["bytes_lost"]
            assert got["rtt_samples"] == rec["rtt_samples"] and got["utility"] == rec["utility"]
            assert got["send_end_time"] == rec["send_end_time"]
    finally:
        th.join(timeout=10)
        link.close()
    assert seen == [str(2.0 + k) for k in range(len(recs))]      # the rate travels as str(float), shim_env.py:107


@pytest.mark.gpu
@pytest.mark.timeout(120)
def test_shim_env_vs_oracle_history_and_reference_rate_formula():
    """60 MIs through the socket, a reset in the middle: observations == the oracle's SenderHistory, rates == the
    reference's apply_action / set_rate (shim_env.py:82-96), reward and done as the reference returns them."""
    import pcc_rl_b200
    g = np.random.default_rng(5)
    recs = _records(g, 60, 200)
    socket.setdefaulttimeout(30)
    env = pcc_rl_b200.ShimNetworkEnv(port=0)
    assert env.observation_space.shape == (30,) and env.action_space.shape == (1,)
    seen = []                     # a test must fail, not wait, if the other side dies
    th = threading.Thread(target=_sender, args=(env.link.port, recs, seen), daemon=True)
    th.start()
    orc = oracle.OracleFlows(1)
    try:
        obs0 = env.reset()
        orc.reset(0, 2)
        assert np.array_equal(obs0, orc.obs(0))
        rate = 2.0
        for k, rec in enumerate(recs):
            if k == 30:
                obs0 = env.reset()
                orc.reset(0, 2)
                rate = 2.0
                assert np.array_equal(obs0, orc.obs(0)) and env.steps_taken == 0
            a = float(g.normal(0, 20.0))
            delta = a * 0.025                                              # shim_env.py:82-96
            rate = rate * (1.0 + delta) if delta >= 0.0 else rate / (1.0 - delta)
            rate = min(max(rate, 0.25), 1000.0)
            obs, rew, done, info = env.step(np.array([a]))
            orc.give_sample(0, rec["bytes_sent"], rec["bytes_acked"], rec["bytes_lost"], rec["send_start_time"],
                            rec["send_end_time"], rec["recv_start_time"], rec["recv_end_time"], rec["rtt_samples"],
                            rec["packet_size"])
            assert env.rate == rate, k
            assert np.array_equal(obs, orc.obs(0)), k
            assert rew == rec["utility"] and done is False and info == {}
        assert len(seen) == 60 and float(seen[-1]) == rate
        env.mon.check()
    finally:
        socket.setdefaulttimeout(None)
        th.join(timeout=10)
        env.close()
