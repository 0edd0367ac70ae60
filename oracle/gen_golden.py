"""Generates tests/golden/*.npz by EXECUTING THE UNMODIFIED REFERENCE -- TEST INFRASTRUCTURE.

Run in the build container (needs /root/reference):   python oracle/gen_golden.py   (all of round 1's fixtures)
                                                      python oracle/gen_golden.py queue1   (round 2: one-packet queues)
The fixtures are data (inputs + the reference's outputs); no reference source is copied.

Two families:
  mt_seed<S>.npz      the reference exactly as shipped: random.seed(S); SimulatedNetworkEnv();
                      episodes of reset() + 400 step()s with N(0,1) actions (Python floats).
                      Link parameters are whatever the reference drew from its global stream.
  philox_<name>.npz   the reference with `network_sim.random` replaced by a per-env Philox
                      stream (oracle/philox_py.py) and scripted link parameters, covering edge
                      cases (no loss, tiny/huge queue, rate >> bw, rate clamps, long/short
                      history, all 12 features).
Per step the fixture stores: action, obs, reward, done, sent/acked/lost, cur_time, run_dur,
rate, and the event-log fields of network_sim.py:422-436.
"""
import math
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refharness as rh  # noqa: E402
from philox_py import PhiloxStream  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
ALL_FEATURES = ("send rate,recv rate,recv dur,send dur,avg latency,loss ratio,"
                "ack latency inflation,sent latency inflation,conn min latency,"
                "latency increase,latency ratio,send ratio")
DEFAULT_FEATURES = "sent latency inflation,latency ratio,send ratio"


def _record_step(env, action, obs, reward, done, rec):
    s = env.senders[0]
    ev = env.event_record["Events"][-1]
    rec["action"].append(action)
    rec["obs"].append(np.asarray(obs, dtype=np.float64))
    rec["reward"].append(float(reward))
    rec["done"].append(bool(done))
    rec["counts"].append((s.sent, s.acked, s.lost))
    rec["cur_time"].append(env.net.cur_time)
    rec["run_dur"].append(float(env.run_dur))
    rec["rate"].append(float(s.rate))
    rec["info"].append((ev["Send Rate"], ev["Throughput"], ev["Latency"], ev["Loss Rate"],
                        ev["Latency Inflation"], ev["Latency Ratio"], ev["Send Ratio"]))


def _new_rec():
    return {k: [] for k in ("action", "obs", "reward", "done", "counts", "cur_time", "run_dur",
                            "rate", "info", "ep_params", "ep_obs0", "ep_cur_time0")}


def _record_reset(env, obs0, rec):
    l, s = env.links[0], env.senders[0]
    queue = int(round(l.max_queue_delay * l.bw))
    rec["ep_params"].append((l.bw, l.dl, float(queue), l.lr, s.starting_rate))
    rec["ep_obs0"].append(np.asarray(obs0, dtype=np.float64))
    rec["ep_cur_time0"].append(env.net.cur_time)


def _save(name, rec, **meta):
    arrs = dict(
        action=np.array(rec["action"]), obs=np.array(rec["obs"]), reward=np.array(rec["reward"]),
        done=np.array(rec["done"]), counts=np.array(rec["counts"], dtype=np.int64),
        cur_time=np.array(rec["cur_time"]), run_dur=np.array(rec["run_dur"]),
        rate=np.array(rec["rate"]), info=np.array(rec["info"]),
        ep_params=np.array(rec["ep_params"]), ep_obs0=np.array(rec["ep_obs0"]),
        ep_cur_time0=np.array(rec["ep_cur_time0"]))
    for k, v in meta.items():
        arrs[k] = np.array(v)
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrs)
    print("wrote %s: %d steps" % (name, len(rec["action"])))


def gen_mt(ns, seed, n_eps, action_seed, zero_actions=False):
    rec = _new_rec()
    arng = random.Random(action_seed)
    with rh.quiet_tmp_cwd():
        random.seed(seed)
        env = ns.SimulatedNetworkEnv()
        for _ in range(n_eps):
            obs0 = env.reset()
            _record_reset(env, obs0, rec)
            done = False
            while not done:
                a = 0.0 if zero_actions else arng.gauss(0.0, 1.0)
                obs, r, done, _ = env.step([a])
                _record_step(env, a, obs, r, done, rec)
    _save("mt_seed%d" % seed, rec, seed=seed, rng="mt19937", history_len=10,
          features=DEFAULT_FEATURES, steps_per_episode=400)


def gen_philox(ns, name, seed, params, n_steps, actions, history_len=10,
               features=DEFAULT_FEATURES, n_eps=1):
    """params: list (one per episode) of (bw, lat, queue, loss, start_factor)."""
    rec = _new_rec()
    stream = PhiloxStream(seed)
    shim = rh.StreamShim([stream])
    real_random = ns.random
    ns.random = shim
    try:
        with rh.quiet_tmp_cwd():
            shim.script = [100.0, 0.1, 0.0, 0.0, 1.0]  # __init__ builds a throw-away link
            env = ns.SimulatedNetworkEnv(history_len=history_len, features=features)
            k = 0
            for ep in range(n_eps):
                bw, lat, queue, loss, factor = params[ep]
                shim.script = [bw, lat, math.log(queue - 1 + 0.5), loss, factor]
                obs0 = env.reset()
                assert not shim.script
                _record_reset(env, obs0, rec)
                assert rec["ep_params"][-1][2] == queue, (rec["ep_params"][-1], queue)
                for t in range(n_steps):
                    a = float(actions[k]); k += 1
                    obs, r, done, _ = env.step([a])
                    _record_step(env, a, obs, r, done, rec)
    finally:
        ns.random = real_random
    _save("philox_" + name, rec, seed=seed, rng="philox", history_len=history_len,
          features=features, steps_per_episode=n_steps)


def gen_multi(ns, name, seed, bw, lat, queue, loss, rates, n_steps, actions):
    """Several senders on one bottleneck: the reference's Network/Link/Sender classes with the external
    patch Sender.__lt__ = id order and the env glue of network_sim.py:406-484 per sender."""
    S = len(rates)
    feats = DEFAULT_FEATURES.split(",")
    real_random, had_lt = ns.random, getattr(ns.Sender, "__lt__", None)
    ns.Sender.__lt__ = lambda a, b: a.id < b.id
    ns.random = rh.StreamShim([PhiloxStream(seed)])
    try:
        with rh.quiet_tmp_cwd():
            links = [ns.Link(bw, lat, queue, loss), ns.Link(bw, lat, queue, loss)]
            senders = [ns.Sender(r, [links[0], links[1]], 0, feats, history_len=10) for r in rates]
            run_dur = 3 * lat
            net = ns.Network(senders, links)
            net.run_for_dur(run_dur)
            net.run_for_dur(run_dur)
            obs_l, rew_l, cnt_l, ct_l, rd_l = [], [], [], [], []
            cur0 = net.cur_time
            for t in range(n_steps):
                for s_, a in zip(senders, actions[t]):
                    s_.apply_rate_delta(float(a))
                net.run_for_dur(run_dur)
                o_t, r_t, c_t = [], [], []
                for i, s_ in enumerate(senders):
                    s_.record_run()
                    o_t.append(np.array(s_.get_obs()).reshape(-1))
                    mi = s_.get_run_data()
                    r_t.append((10.0 * mi.get("recv rate") / (8 * 1500) - 1e3 * mi.get("avg latency")
                                - 2e3 * mi.get("loss ratio")) * 0.001)
                    if i == 0:
                        avg0 = mi.get("avg latency")
                    mi.get("latency ratio")
                    c_t.append((s_.sent, s_.acked, s_.lost))
                if avg0 > 0.0:
                    run_dur = 0.5 * avg0
                obs_l.append(o_t); rew_l.append(r_t); cnt_l.append(c_t); ct_l.append(net.cur_time); rd_l.append(run_dur)
    finally:
        ns.random = real_random
        if had_lt is None:
            del ns.Sender.__lt__
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "multi_" + name + ".npz"), seed=np.array(seed), rng=np.array("philox"),
                        params=np.array([bw, lat, float(queue), loss]), rates=np.array(rates),
                        action=np.array(actions), obs=np.array(obs_l), reward=np.array(rew_l),
                        counts=np.array(cnt_l, dtype=np.int64), cur_time=np.array(ct_l), run_dur=np.array(rd_l),
                        cur_time0=np.array(cur0))
    print("wrote multi_%s: %d steps x %d senders" % (name, n_steps, S))


def main():
    ns = rh.load_reference()
    # --- the reference exactly as shipped (global MT19937 stream) ---
    gen_mt(ns, 1234, 2, 7)            # the KAT seed of SURVEY.md §8a
    gen_mt(ns, 100, 2, 101)           # BASELINE.md's benchmark seeds
    gen_mt(ns, 2019, 1, 3, zero_actions=True)
    gen_mt(ns, 987654321987, 1, 5)    # a seed wider than 32 bits (two init_by_array limbs)
    # --- scripted edge cases on Philox streams ---
    g = random.Random(42)
    acts = lambda n, s=1.0: [g.gauss(0.0, s) for _ in range(n)]
    gen_philox(ns, "noloss", 11, [(300.0, 0.1, 50, 0.0, 0.5)], 120, acts(120))
    gen_philox(ns, "tinyqueue_overdrive", 12, [(100.0, 0.05, 2, 0.01, 1.5)], 120, [4.0] * 60 + acts(60, 3.0))
    gen_philox(ns, "hugequeue_bufferbloat", 13, [(100.0, 0.2, 2981, 0.0, 1.5)], 100, [6.0] * 40 + [-6.0] * 60)
    gen_philox(ns, "maxrate_clamp", 14, [(500.0, 0.05, 10, 0.05, 1.5)], 100, [20.0] * 100)
    gen_philox(ns, "minrate_clamp", 15, [(100.0, 0.5, 5, 0.02, 0.3)], 80, [-20.0] * 80)
    gen_philox(ns, "heavyloss", 16, [(200.0, 0.08, 20, 0.6, 1.0)], 100, acts(100, 2.0))
    gen_philox(ns, "allfeatures", 17, [(250.0, 0.12, 8, 0.03, 1.2), (120.0, 0.3, 100, 0.0, 0.9)], 100,
               acts(200, 2.0), features=ALL_FEATURES, n_eps=2)
    gen_philox(ns, "hist1", 18, [(400.0, 0.06, 3, 0.04, 1.4)], 60, acts(60, 2.0), history_len=1)
    gen_philox(ns, "hist25_rates", 19, [(150.0, 0.25, 30, 0.01, 0.7)], 60, acts(60, 2.0), history_len=25,
               features="send rate,recv rate,avg latency,loss ratio")
    gen_philox(ns, "highbw_idle", 20, [(83333.0, 0.001, 1000, 0.0, 0.01)], 100, acts(100, 2.0))
    gen_philox(ns, "bigseed", 0xFEDCBA9876543210, [(333.0, 0.07, 6, 0.05, 1.3)], 60, acts(60, 2.0))
    # --- BASELINE config 5: two (three) senders per link, points of the bw x delay grid ---
    macts = lambda n, S, s=2.0: [[g.gauss(0.0, s) for _ in range(S)] for _ in range(n)]
    gen_multi(ns, "2s_lowbw", 31, 83.3, 0.05, 10, 0.0, [120.0, 60.0], 100, macts(100, 2))
    gen_multi(ns, "2s_midbw_loss", 32, 833.0, 0.02, 30, 0.02, [500.0, 700.0], 100, macts(100, 2))
    gen_multi(ns, "2s_highbw_shortlat", 33, 83333.0, 0.001, 100, 0.0, [900.0, 40.0], 100, macts(100, 2))
    gen_multi(ns, "2s_longlat", 34, 400.0, 0.5, 5, 0.01, [300.0, 300.0], 60, macts(60, 2))
    gen_multi(ns, "3s_tinyqueue", 35, 200.0, 0.03, 2, 0.05, [150.0, 150.0, 150.0], 80, macts(80, 3))


def main_queue1():
    """Round 2: a queue of exactly ONE packet (max_queue_delay == 1/bw) -- legal for the reference's classes (its sampler
    draws 1 + int(exp(U(0, 8))) >= 2, the scripted stream below makes int(exp(.)) = 0), and the corner where the tail-drop
    threshold of pcc_core.cuh sits within one ulp of w = 0.  Own action stream: the fixtures above stay byte-identical."""
    ns = rh.load_reference()
    g = random.Random(43)
    acts = lambda n, s=1.0: [g.gauss(0.0, s) for _ in range(n)]
    gen_philox(ns, "queue1_overdrive", 41, [(150.0, 0.05, 1, 0.01, 1.5)], 100, [3.0] * 40 + acts(60, 3.0))
    gen_philox(ns, "queue1_highbw_shortlat", 42, [(1054.5534236464064, 0.0023996703096765487, 1, 0.05, 0.72)], 80,
               acts(80, 2.0))
    macts = lambda n, S, s=2.0: [[g.gauss(0.0, s) for _ in range(S)] for _ in range(n)]
    gen_multi(ns, "2s_queue1", 43, 300.0, 0.04, 1, 0.02, [200.0, 350.0], 80, macts(80, 2))


if __name__ == "__main__":
    main_queue1() if sys.argv[1:] == ["queue1"] else main()
