"""Generates tests/golden/variant_*.npz by EXECUTING THE UNMODIFIED REFERENCE with its own module switches
turned on -- TEST INFRASTRUCTURE (needs /root/reference).   python oracle/gen_golden_variants.py

network_sim.py ships with USE_CWND = False and USE_LATENCY_NOISE = False (:51-54); the code paths they guard
(:150-151, 171-172 latency noise; :243-255, 283-289, 413-414 congestion window, 2-dim action) are part of the
same event loop.  The switches are module globals read at call time, so they are set on
the imported module object (`ns.USE_CWND = True`) -- no reference file is modified -- and the env is driven as
usual with `network_sim.random` replaced by a Philox stream (oracle/philox_py.py), like the philox_* goldens.
Multi-sender files additionally use the documented Sender.__lt__ patch.
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
import gen_golden as gg  # noqa: E402

OUT = gg.OUT


def gen_variant(ns, name, seed, params, n_steps, actions, cwnd_actions, use_cwnd, use_noise, features=gg.DEFAULT_FEATURES):
    rec = gg._new_rec()
    rec["cwnd"] = []
    shim = rh.StreamShim([PhiloxStream(seed)])
    real_random, old = ns.random, (ns.USE_CWND, ns.USE_LATENCY_NOISE)
    ns.random = shim
    ns.USE_CWND, ns.USE_LATENCY_NOISE = use_cwnd, use_noise
    try:
        with rh.quiet_tmp_cwd():
            shim.script = [100.0, 0.1, 0.0, 0.0, 1.0]
            env = ns.SimulatedNetworkEnv(history_len=10, features=features)
            assert env.action_space.shape == ((2,) if use_cwnd else (1,))
            k = 0
            for ep in range(len(params)):
                bw, lat, queue, loss, factor = params[ep]
                shim.script = [bw, lat, math.log(queue - 1 + 0.5), loss, factor]
                obs0 = env.reset()
                assert not shim.script
                gg._record_reset(env, obs0, rec)
                for t in range(n_steps):
                    a, c = float(actions[k]), float(cwnd_actions[k])
                    k += 1
                    obs, r, done, _ = env.step([a, c] if use_cwnd else [a])
                    gg._record_step(env, a, obs, r, done, rec)
                    rec["cwnd"].append(int(env.senders[0].cwnd))
    finally:
        ns.random = real_random
        ns.USE_CWND, ns.USE_LATENCY_NOISE = old
    gg._save("variant_" + name, rec, seed=seed, rng="philox", history_len=10, features=features,
             steps_per_episode=n_steps, use_cwnd=use_cwnd, use_noise=use_noise,
             cwnd=np.array(rec["cwnd"], dtype=np.int64), cwnd_action=np.array(cwnd_actions[:len(rec["action"])]))


def gen_multi_variant(ns, name, seed, bw, lat, queue, loss, rates, n_steps, actions, cwnd_actions, use_cwnd, use_noise):
    """gen_golden.gen_multi with the switches on and per-sender cwnd actions."""
    S = len(rates)
    feats = gg.DEFAULT_FEATURES.split(",")
    real_random, had_lt, old = ns.random, getattr(ns.Sender, "__lt__", None), (ns.USE_CWND, ns.USE_LATENCY_NOISE)
    ns.Sender.__lt__ = lambda a, b: a.id < b.id
    ns.random = rh.StreamShim([PhiloxStream(seed)])
    ns.USE_CWND, ns.USE_LATENCY_NOISE = use_cwnd, use_noise
    try:
        with rh.quiet_tmp_cwd():
            links = [ns.Link(bw, lat, queue, loss), ns.Link(bw, lat, queue, loss)]
            senders = [ns.Sender(r, [links[0], links[1]], 0, feats, history_len=10) for r in rates]
            run_dur = 3 * lat
            net = ns.Network(senders, links)
            net.run_for_dur(run_dur)
            net.run_for_dur(run_dur)
            obs_l, rew_l, cnt_l, ct_l, rd_l, cw_l = [], [], [], [], [], []
            cur0 = net.cur_time
            for t in range(n_steps):
                for i, s_ in enumerate(senders):
                    s_.apply_rate_delta(float(actions[t][i]))
                    if use_cwnd:
                        s_.apply_cwnd_delta(float(cwnd_actions[t][i]))
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
                cw_l.append([int(s_.cwnd) for s_ in senders])
    finally:
        ns.random = real_random
        ns.USE_CWND, ns.USE_LATENCY_NOISE = old
        if had_lt is None:
            del ns.Sender.__lt__
    np.savez_compressed(os.path.join(OUT, "variant_multi_" + name + ".npz"), seed=np.array(seed), rng=np.array("philox"),
                        params=np.array([bw, lat, float(queue), loss]), rates=np.array(rates),
                        action=np.array(actions), cwnd_action=np.array(cwnd_actions), obs=np.array(obs_l),
                        reward=np.array(rew_l), counts=np.array(cnt_l, dtype=np.int64), cur_time=np.array(ct_l),
                        run_dur=np.array(rd_l), cur_time0=np.array(cur0), cwnd=np.array(cw_l, dtype=np.int64),
                        use_cwnd=np.array(use_cwnd), use_noise=np.array(use_noise))
    print("wrote variant_multi_%s: %d steps x %d senders" % (name, n_steps, S))


def main():
    ns = rh.load_reference()
    g = random.Random(4242)
    acts = lambda n, s=1.0: [g.gauss(0.0, s) for _ in range(n)]
    # congestion window only: the window starts at 25 and binds whenever rate x RTT > cwnd
    gen_variant(ns, "cwnd", 51, [(300.0, 0.1, 50, 0.01, 1.2), (150.0, 0.3, 10, 0.0, 1.5)], 100, acts(200, 2.0),
                acts(200, 4.0), True, False)
    gen_variant(ns, "cwnd_shrink_to_min", 52, [(400.0, 0.2, 20, 0.02, 1.5)], 120, [3.0] * 120, [-8.0] * 120, True, False)
    gen_variant(ns, "cwnd_grow_to_max", 53, [(500.0, 0.05, 100, 0.0, 1.0)], 150, acts(150, 1.0), [40.0] * 150, True, False)
    # latency noise only: one extra uniform draw per hop
    gen_variant(ns, "noise", 54, [(250.0, 0.12, 8, 0.03, 1.2), (120.0, 0.3, 100, 0.0, 0.9)], 100, acts(200, 2.0),
                [0.0] * 200, False, True, features=gg.ALL_FEATURES)
    gen_variant(ns, "noise_overdrive", 55, [(100.0, 0.05, 2, 0.01, 1.5)], 100, [4.0] * 50 + acts(50, 3.0), [0.0] * 100,
                False, True)
    # both
    gen_variant(ns, "cwnd_noise", 56, [(200.0, 0.08, 20, 0.05, 1.0), (333.0, 0.07, 6, 0.05, 1.3)], 100, acts(200, 2.0),
                acts(200, 3.0), True, True)
    macts = lambda n, S, s=2.0: [[g.gauss(0.0, s) for _ in range(S)] for _ in range(n)]
    gen_multi_variant(ns, "2s_cwnd_noise", 57, 400.0, 0.05, 10, 0.01, [300.0, 500.0], 80, macts(80, 2), macts(80, 2, 4.0),
                      True, True)
    gen_multi_variant(ns, "2s_cwnd", 58, 833.0, 0.02, 30, 0.0, [900.0, 700.0], 80, macts(80, 2), macts(80, 2, 4.0),
                      True, False)


if __name__ == "__main__":
    main()
