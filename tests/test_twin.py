"""The product's per-env MI code (pcc_core.cuh), compiled for the host, against (a) the golden
outputs of the unmodified reference and (b) the heap-based oracle on randomized episodes.
This is where the three-cursor streaming algorithm and its tie handling are proven before
the same header is compiled into the CUDA kernels."""
import numpy as np
import pytest

import oracle
from golden_util import assert_step_equal, golden_names, load_golden
from twin_util import TwinEnv


@pytest.mark.parametrize("name", golden_names("mt_"))
def test_twin_matches_reference_as_shipped(name):
    g = load_golden(name)
    o = oracle.OracleEnv()  # only as the host-side MT parameter sampler (random.uniform order)
    o.seed_mt(g["seed"])
    o.sample_params()
    t = TwinEnv(g["history_len"], g["features"])
    k = 0
    for ep in range(len(g["ep_params"])):
        bw, lat, q, loss, rate = o.sample_params()
        assert (bw, lat, float(q), loss, rate) == tuple(g["ep_params"][ep])
        t.mt_setstate(o.mt_getstate())
        obs0 = t.reset(bw, lat, q, loss, rate)
        assert np.array_equal(obs0, g["ep_obs0"][ep]) and t.cur_time == g["ep_cur_time0"][ep]
        for _ in range(g["steps_per_episode"]):
            obs, r, d, c, info = t.step(g["action"][k])
            assert_step_equal(g, k, obs, r, d, c, t.cur_time, t.run_dur, t.rate, info, name)
            k += 1
        o.mt_setstate(t.mt_getstate())
    assert not t.overflow


@pytest.mark.parametrize("name", golden_names("philox_"))
def test_twin_matches_reference_on_philox_stream(name):
    g = load_golden(name)
    t = TwinEnv(g["history_len"], g["features"])
    t.seed_philox(g["seed"])
    k = 0
    for ep in range(len(g["ep_params"])):
        bw, lat, q, loss, rate = g["ep_params"][ep]
        obs0 = t.reset(bw, lat, int(q), loss, rate)
        assert np.array_equal(obs0, g["ep_obs0"][ep]) and t.cur_time == g["ep_cur_time0"][ep]
        for _ in range(g["steps_per_episode"]):
            obs, r, d, c, info = t.step(g["action"][k])
            assert_step_equal(g, k, obs, r, d, c, t.cur_time, t.run_dur, t.rate, info, name)
            k += 1
    assert not t.overflow


def _lockstep(seed, n_eps, n_steps, sampler, act_sigma, ring_capacity=1 << 16, ring_base=0,
              features=oracle.DEFAULT_FEATURES, prescan=False):
    g = np.random.default_rng(seed)
    o = oracle.OracleEnv(10, features)
    t = TwinEnv(10, features, ring_capacity)
    o.seed_philox(seed)
    t.seed_philox(seed)
    t.set_ring_cursor(ring_base)
    t.set_prescan(prescan)
    steps = 0
    for ep in range(n_eps):
        p = sampler(g)
        a0 = o.reset(*p)
        b0 = t.reset(*p)
        assert np.array_equal(a0, b0) and o.cur_time == t.cur_time, (seed, ep)
        for k in range(n_steps):
            a = float(g.normal(0, act_sigma)) if act_sigma > 0 else 0.0
            x = o.step(a)
            y = t.step(a)
            ctx = (seed, ep, k, p)
            assert tuple(x[3]) == tuple(y[3]), ctx
            assert np.array_equal(x[0], y[0]), ctx
            assert x[1] == y[1] and x[2] == y[2], ctx
            assert np.array_equal(x[4], y[4]), ctx
            assert o.cur_time == t.cur_time and o.run_dur == t.run_dur and o.rate == t.rate, ctx
            steps += 1
    assert not t.overflow
    return steps


def default_ranges(g):
    """create_new_links_and_senders, network_sim.py:455-466 (ICML'19 default ranges)."""
    bw = g.uniform(100, 500)
    return (bw, g.uniform(0.05, 0.5), 1 + int(np.exp(g.uniform(0, 8))), g.uniform(0, 0.05),
            g.uniform(0.3, 1.5) * bw)


def nasty_ranges(g):
    """Heavier loss, overdriven tiny queues, wide bandwidths: maximises ties and clusters."""
    bw = float(np.exp(g.uniform(np.log(40), np.log(5000))))
    return (bw, float(np.exp(g.uniform(np.log(0.001), np.log(0.5)))), 1 + int(np.exp(g.uniform(0, 5))),
            float(g.choice([0.0, 0.01, 0.3, 0.9, 1.0])), float(g.uniform(40, 1000)))


@pytest.mark.parametrize("seed", range(12))
def test_twin_equals_oracle_default_ranges(seed):
    _lockstep(1000 + seed, n_eps=3, n_steps=400, sampler=default_ranges, act_sigma=1.0)


@pytest.mark.parametrize("seed", range(6))
def test_twin_two_stage_consumption_equals_oracle(seed):
    """The claim behind the helper warp of the small-batch step kernel: scanning the cursors over the records that
    exist before the MI's sends, then resuming with the final tail, changes nothing (prefix scans)."""
    _lockstep(3000 + seed, n_eps=2, n_steps=400, sampler=default_ranges, act_sigma=1.0, prescan=True)
    _lockstep(4000 + seed, n_eps=3, n_steps=150, sampler=nasty_ranges, act_sigma=3.0, prescan=True,
              features="send rate,recv rate,avg latency,loss ratio,sent latency inflation,conn min latency,"
                       "latency increase,latency ratio,send ratio")


@pytest.mark.parametrize("seed", range(12))
def test_twin_equals_oracle_nasty_ranges(seed):
    _lockstep(2000 + seed, n_eps=4, n_steps=150, sampler=nasty_ranges, act_sigma=3.0,
              features="send rate,recv rate,recv dur,send dur,avg latency,loss ratio,"
                       "ack latency inflation,sent latency inflation,conn min latency,"
                       "latency increase,latency ratio,send ratio")


def tiny_queues(g):
    """Queues of 0, 1 and 2 packets (the reference's own sampler never goes below 2, its classes take any): with one
    packet max_queue_delay == 1/bw, so the tail-drop threshold is the ulp-sized interval around w = 0."""
    bw = float(np.exp(g.uniform(np.log(40), np.log(5000))))
    return (bw, float(np.exp(g.uniform(np.log(0.001), np.log(0.5)))), int(g.integers(0, 3)),
            float(g.choice([0.0, 0.05, 0.3, 1.0])), float(g.uniform(40, 1000)))


@pytest.mark.parametrize("seed", range(6))
def test_twin_equals_oracle_tiny_queues(seed):
    _lockstep(5000 + seed, n_eps=6, n_steps=100, sampler=tiny_queues, act_sigma=3.0)


def test_twin_ring_wraps_u32_and_small_capacity():
    """Ring positions are u32 counters: run across the 2^32 wrap with a ring of 4096 slots."""
    small = lambda g: (g.uniform(100, 300), g.uniform(0.05, 0.2), 1 + int(np.exp(g.uniform(0, 4))),
                       g.uniform(0, 0.05), g.uniform(40, 300))
    _lockstep(7, n_eps=3, n_steps=400, sampler=small, act_sigma=1.0, ring_capacity=4096,
              ring_base=2**32 - 20000)


def test_twin_reports_ring_overflow():
    t = TwinEnv(ring_capacity=64)
    t.seed_philox(1)
    t.reset(500.0, 0.5, 1000, 0.0, 750.0)
    assert t.overflow


def test_tail_drop_threshold_is_exact():
    """pcc_core.cuh: tail_drop_threshold -- `w > w_full` must equal the reference's test
    `1/bw + w > max_queue_delay` (network_sim.py:77-79) for EVERY w >= 0, in particular in the ulp
    neighbourhood of the threshold."""
    import ctypes as C
    import twin_util
    L = twin_util.lib()
    L.twin_tail_drop_threshold.restype = C.c_double
    L.twin_tail_drop_threshold.argtypes = [C.c_double, C.c_double]
    g = np.random.default_rng(0)
    for _ in range(3000):
        bw = float(np.exp(g.uniform(np.log(1.0), np.log(1e6))))
        queue = int(1 + np.exp(g.uniform(0, 9)))
        d_bw, max_qd = 1.0 / bw, queue / bw
        wf = L.twin_tail_drop_threshold(d_bw, max_qd)
        assert wf >= 0.0
        cands = [0.0, wf, max_qd, max_qd - d_bw, g.uniform(0, 2 * max_qd)]
        w = wf
        for _k in range(4):
            w = np.nextafter(w, np.inf); cands.append(float(w))
        w = wf
        for _k in range(4):
            w = np.nextafter(w, -np.inf)
            if w >= 0:
                cands.append(float(w))
        for w in cands:
            assert (d_bw + w > max_qd) == (w > wf), (bw, queue, w, wf)
    assert L.twin_tail_drop_threshold(1.0, 0.5) == -1.0   # never admissible: every packet is tail-dropped
    # a queue of exactly one packet: max_qd == d_bw, the admissible interval is [0, w_full] with w_full below one ulp
    # of d_bw (walking there ulp by ulp from 0 would never end)
    for bw in (83.3, 1054.5534236464064, 1e5):
        d_bw = 1.0 / bw
        wf = L.twin_tail_drop_threshold(d_bw, 1.0 / bw)
        assert 0.0 <= wf < np.spacing(d_bw)
        assert d_bw + wf <= d_bw and d_bw + float(np.nextafter(wf, np.inf)) > d_bw


def test_pw_stream_push_form_equals_pull_form():
    """PwStream (the push-style pairwise sum of the lane-per-env kernels) == np_mean_stream / pw_sum (pinned to numpy
    by tests/test_oracle_golden.py) for every sample count 1..3000: total, first half and second half."""
    import ctypes as C
    import twin_util
    L = twin_util.lib()
    L.twin_pw_stream_mismatches.argtypes = [C.POINTER(C.c_double), C.c_int]
    rng = np.random.default_rng(5)
    for scale in (1.0, 1e-3):
        a = np.ascontiguousarray(rng.uniform(0.05, 1.3, 3000) * scale)
        assert L.twin_pw_stream_mismatches(a.ctypes.data_as(C.POINTER(C.c_double)), 3000) == 0
    # and against numpy itself
    a = rng.uniform(0.05, 1.3, 20000)
    for n in (1, 7, 8, 9, 127, 128, 129, 255, 256, 257, 1000, 4097, 20000):
        assert L.twin_pw_stream_mismatches(a.ctypes.data_as(C.POINTER(C.c_double)), n) == 0


def test_integer_loss_draw_equals_double_compare():
    """loss_threshold / u53 (the integer form of `random.random() < lr` used by the packed kernels): same decision as
    the binary64 compare for random and adversarial loss rates, incl. draws that land exactly on / next to the rate."""
    import ctypes as C
    import twin_util
    L = twin_util.lib()
    L.twin_loss_threshold_mismatches.restype = C.c_long
    L.twin_loss_threshold_mismatches.argtypes = [C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_uint32),
                                                 C.POINTER(C.c_uint32), C.c_int]
    rng = np.random.default_rng(9)
    a = rng.integers(0, 2**32, 4000, dtype=np.uint64).astype(np.uint32)
    b = rng.integers(0, 2**32, 4000, dtype=np.uint64).astype(np.uint32)
    # draws equal to, just below and just above a handful of rates
    k = rng.integers(0, 2**53, 200, dtype=np.uint64)
    exact = (k.astype(np.float64) / 2.0**53)        # exact: k < 2^53
    lr = np.concatenate([rng.uniform(0, 0.05, 300), rng.uniform(0, 1, 100), exact, np.nextafter(exact, 0), np.nextafter(exact, 1),
                         [0.0, -0.0, -1.0, 1.0, 1.5, np.nan, 5e-324, 2.0**-53, 2.0**-54, 1 - 2.0**-53, 0.5]])
    a2 = np.concatenate([a, (k >> np.uint64(26)).astype(np.uint32) << np.uint32(5)])
    b2 = np.concatenate([b, (k & np.uint64((1 << 26) - 1)).astype(np.uint32) << np.uint32(6)])
    lr = np.ascontiguousarray(lr)
    bad = L.twin_loss_threshold_mismatches(lr.ctypes.data_as(C.POINTER(C.c_double)), len(lr),
                                           np.ascontiguousarray(a2).ctypes.data_as(C.POINTER(C.c_uint32)),
                                           np.ascontiguousarray(b2).ctypes.data_as(C.POINTER(C.c_uint32)), len(a2))
    assert bad == 0
