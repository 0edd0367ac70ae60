"""Generates tests/golden/flows_*.npz by RUNNING THE UNMODIFIED REFERENCE in this container
(test infrastructure; needs /root/reference).  python oracle/gen_golden_flows.py

  flows_client_default   udt-plugins/testing/loaded_client.py driven through its module API
                         (init / give_sample / get_rate / reset) with a stub agent (TensorFlow absent),
                         default 3 features, history 10, 16 flows
  flows_allfeatures      common/sender_obs.py directly (SenderHistory.step + as_array), all 12 metrics,
                         history 5, every sample-count edge case of numpy's pairwise mean
  flows_shim             gym/online/shim_env.ShimNetworkEnv.step fed by udt-plugins/training/shim.py
                         over a real localhost socket: the wire format (%f text) and the reset semantics

Each file stores the operation stream (op 0 = record, 1 = history reset) with every input field, the CSR
sample arrays, and what the reference returned after each op (observation, rate, metrics).
"""
import os
import sys
import threading

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refharness  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
ALL_FEATURES = ("send rate,recv rate,recv dur,send dur,avg latency,loss ratio,ack latency inflation,"
                "sent latency inflation,conn min latency,latency increase,latency ratio,send ratio")
EDGE_COUNTS = [0, 1, 2, 3, 7, 8, 9, 15, 16, 17, 63, 64, 65, 127, 128, 129, 130, 255, 256, 257, 271, 272, 273,
               511, 512, 513, 1000, 1023, 1024, 1025, 2047, 2049, 4001]


def make_record(g, t, kind="normal", n=None, quantise=False):
    """One synthetic MI record.  Returns (fields dict, rtt list, end time)."""
    ps = int(g.choice([1500, 1500, 1400, 1000]))
    dur = float(g.uniform(0.01, 0.5))
    base = float(g.uniform(0.01, 0.4))
    if n is None:
        n = int(g.integers(0, 400)) if g.random() < 0.9 else int(g.integers(400, 3000))
    lost = int(g.integers(0, max(1, n // 10 + 1))) if g.random() < 0.6 else 0
    sent = n + lost + int(g.integers(0, 5))
    rtt = base + base * 0.5 * g.random(n) * np.linspace(1.0, float(g.uniform(0.5, 2.0)), n) if n else np.zeros(0)
    f = dict(bytes_sent=sent * ps, bytes_acked=n * ps, bytes_lost=lost * ps, send_start=t, send_end=t + dur,
             recv_start=t + base, recv_end=t + dur + base * float(g.uniform(0.9, 1.5)), packet_size=ps)
    if kind == "zero_send_dur":
        f["send_end"] = f["send_start"]
    elif kind == "zero_recv_dur":
        f["recv_end"] = f["recv_start"]
    elif kind == "negative_dur":
        f["send_end"] = f["send_start"] - 0.001
        f["recv_end"] = f["recv_start"] - 0.002
    elif kind == "nothing_acked":
        f["bytes_acked"] = 0
        rtt = np.zeros(0)
    elif kind == "all_zero":
        f.update(bytes_sent=0, bytes_acked=0, bytes_lost=0)
        rtt = np.zeros(0)
    elif kind == "tiny_throughput":     # send_rate >= 1000 * recv_rate -> send ratio 1.0
        f["bytes_acked"] = ps + 1
        f["bytes_sent"] = 5000 * ps
        rtt = rtt[:1] if len(rtt) else np.array([base])
    elif kind == "odd_bytes":           # byte counts that are not multiples of the packet size
        f["bytes_sent"] += int(g.integers(1, ps))
        f["bytes_acked"] += int(g.integers(1, ps))
        f["bytes_lost"] += int(g.integers(0, ps))
    if quantise:   # what survives the shim's "%f" wire format
        for k in ("send_start", "send_end", "recv_start", "recv_end"):
            f[k] = float("%f" % f[k])
        rtt = np.array([float("%f" % v) for v in rtt])
    return f, [float(v) for v in rtt], f["send_end"]


class Recorder(object):
    def __init__(self, n_flows, history_len, features, reset_mode):
        self.meta = dict(n_flows=n_flows, history_len=history_len, features=features, reset_mode=reset_mode)
        self.rows = []
        self.rtt = []

    def add(self, op, flow, f, rtt, obs, rate=np.nan, action=np.nan, metrics=None):
        self.rows.append((op, flow, f, len(rtt), np.asarray(obs, dtype=np.float64).copy(), rate, action,
                          np.full(12, np.nan) if metrics is None else np.asarray(metrics, dtype=np.float64)))
        self.rtt.extend(rtt)

    def save(self, name):
        E = len(self.rows)
        zero = dict(bytes_sent=0, bytes_acked=0, bytes_lost=0, send_start=0.0, send_end=0.0, recv_start=0.0,
                    recv_end=0.0, packet_size=1500)
        col = lambda k, dt: np.array([(r[2] or zero)[k] for r in self.rows], dtype=dt)
        off = np.zeros(E + 1, dtype=np.int64)
        off[1:] = np.cumsum([r[3] for r in self.rows])
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(
            path, op=np.array([r[0] for r in self.rows], dtype=np.int32),
            flow=np.array([r[1] for r in self.rows], dtype=np.int32),
            bytes_sent=col("bytes_sent", np.int64), bytes_acked=col("bytes_acked", np.int64),
            bytes_lost=col("bytes_lost", np.int64), packet_size=col("packet_size", np.int64),
            send_start=col("send_start", np.float64), send_end=col("send_end", np.float64),
            recv_start=col("recv_start", np.float64), recv_end=col("recv_end", np.float64),
            rtt_off=off, rtt=np.asarray(self.rtt, dtype=np.float64),
            obs=np.stack([r[4] for r in self.rows]), rate=np.array([r[5] for r in self.rows]),
            action=np.array([r[6] for r in self.rows]), metrics=np.stack([r[7] for r in self.rows]),
            **{k: np.array(v) for k, v in self.meta.items()})
        print("%-24s %5d ops, %7d samples -> %s" % (name, E, len(self.rtt), path))


def gen_client_default():
    so, lc = refharness.load_reference_flows()
    g = np.random.default_rng(20260417)
    w = g.normal(0, 0.3, 30)
    refharness.StubAgent.fn = staticmethod(lambda ob: float(2.0 * np.tanh(np.dot(w, np.asarray(ob).reshape(-1)))))
    n_flows = 16
    ids = [1000 + 7 * i for i in range(n_flows)]          # arbitrary flow ids, as the C++ side hands out
    rec = Recorder(n_flows, 10, "sent latency inflation,latency ratio,send ratio", 1)
    for fid in ids:
        lc.init(fid)
    clock = np.zeros(n_flows)
    kinds = ["normal"] * 12 + ["zero_send_dur", "zero_recv_dur", "negative_dur", "nothing_acked", "all_zero",
                               "tiny_throughput", "odd_bytes"]
    for it in range(720):
        i = int(g.integers(0, n_flows))
        if g.random() < 0.04:
            lc.reset(ids[i])                                   # agent.reset + reset_rate + reset_history
            drv = lc.PccGymDriver.get_by_flow_id(ids[i])
            rec.add(1, i, None, [], drv.history.as_array(), rate=drv.rate)
            continue
        f, rtt, clock[i] = make_record(g, clock[i], kind=str(g.choice(kinds)))
        lc.give_sample(ids[i], f["bytes_sent"], f["bytes_acked"], f["bytes_lost"], f["send_start"], f["send_end"],
                       f["recv_start"], f["recv_end"], rtt, f["packet_size"], 0.0)
        drv = lc.PccGymDriver.get_by_flow_id(ids[i])
        obs = drv.history.as_array()
        action = refharness.StubAgent.fn(obs)
        rate = lc.get_rate(ids[i])                             # agent.act(as_array) -> apply_rate_delta; * 1e6
        rec.add(0, i, f, rtt, obs, rate=rate, action=action)
    rec.save("flows_client_default")


def gen_allfeatures():
    so, _ = refharness.load_reference_flows()
    g = np.random.default_rng(99)
    feats = ALL_FEATURES.split(",")
    n_flows, H = 6, 5
    base_id = 770000
    rec = Recorder(n_flows, H, ALL_FEATURES, 1)
    hist = [so.SenderHistory(H, feats, base_id + i) for i in range(n_flows)]
    clock = np.zeros(n_flows)
    plan = [(c, "normal") for c in EDGE_COUNTS] + [(None, k) for k in
            ("zero_send_dur", "zero_recv_dur", "negative_dur", "nothing_acked", "all_zero", "tiny_throughput",
             "odd_bytes")] * 3 + [(None, "normal")] * 60
    order = g.permutation(len(plan))
    for step, pi in enumerate(order):
        n, kind = plan[pi]
        i = int(g.integers(0, n_flows))
        if step % 37 == 36:
            hist[i] = so.SenderHistory(H, feats, base_id + i)  # PccGymDriver.reset_history semantics
            rec.add(1, i, None, [], hist[i].as_array())
            continue
        f, rtt, clock[i] = make_record(g, clock[i], kind=kind, n=n)
        mi = so.SenderMonitorInterval(base_id + i, bytes_sent=f["bytes_sent"], bytes_acked=f["bytes_acked"],
                                      bytes_lost=f["bytes_lost"], send_start=f["send_start"], send_end=f["send_end"],
                                      recv_start=f["recv_start"], recv_end=f["recv_end"], rtt_samples=rtt,
                                      packet_size=f["packet_size"])
        hist[i].step(mi)
        obs = hist[i].as_array()
        rec.add(0, i, f, rtt, obs, metrics=[mi.get(name) for name in feats])
    rec.save("flows_allfeatures")


def gen_shim():
    """ShimNetworkEnv.step (server) <- localhost:9787 <- PccShimDriver.give_sample (client), both unmodified."""
    shim_env, shim = refharness.load_reference_shim()
    g = np.random.default_rng(4242)
    flow_id = 31
    with refharness.quiet_tmp_cwd():
        env = shim_env.ShimNetworkEnv()
    rec = Recorder(1, 10, "sent latency inflation,latency ratio,send ratio", 2)
    T = 90
    records, t = [], 0.0
    for k in range(T):
        n = int(g.integers(0, 24))
        kind = str(g.choice(["normal"] * 6 + ["zero_send_dur", "nothing_acked", "odd_bytes"]))
        f, rtt, t = make_record(g, t, kind=kind, n=n, quantise=True)
        records.append((f, rtt, float("%f" % g.normal(0, 5))))
    actions = g.normal(0, 1, T)

    def client():   # (no stdout redirect here: the main thread's quiet_tmp_cwd already swaps sys.stdout)
        shim.init(flow_id)
        for f, rtt, util in records:
            shim.get_rate(flow_id)
            shim.give_sample(flow_id, f["bytes_sent"], f["bytes_acked"], f["bytes_lost"], f["send_start"],
                             f["send_end"], f["recv_start"], f["recv_end"], rtt, f["packet_size"], util)

    env.sock.listen()                      # so that the client's connect() cannot race the first step()
    th = threading.Thread(target=client, daemon=True)
    th.start()
    with refharness.quiet_tmp_cwd():
        for k in range(T):
            if k == 40:
                obs = env.reset()
                rec.add(1, 0, None, [], obs, rate=env.rate)
            obs, rew, done, _ = env.step([float(actions[k])])   # Python float: keeps the arithmetic binary64 under numpy 2
            f, rtt, util = records[k]
            assert rew == util
            rec.add(0, 0, f, rtt, obs, rate=env.rate, action=float(actions[k]))
    th.join(10)
    env.sock.close()
    if env.conn is not None:
        env.conn.close()
    rec.meta["flow_id"] = flow_id
    rec.save("flows_shim")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    gen_client_default()
    gen_allfeatures()
    gen_shim()
