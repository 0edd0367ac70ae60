"""pcc_multi_core.cuh (several senders on one bottleneck, the product's generic heap path) compiled for the host,
against the reference's golden outputs and against the oracle on random grid-sweep points."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle
import twin_util
from golden_util import GOLDEN_DIR, golden_names


def _lib():
    L = twin_util.lib()
    vp, d, i = C.c_void_p, C.c_double, C.c_int
    pd = C.POINTER(C.c_double)
    L.twin_multi_create.restype = vp
    L.twin_multi_create.argtypes = [i, i, C.POINTER(C.c_int), i, i]
    L.twin_multi_destroy.argtypes = [vp]
    L.twin_multi_seed.argtypes = [vp, C.c_uint64]
    L.twin_multi_reset.argtypes = [vp, d, d, C.c_int64, d, pd]
    L.twin_multi_step.argtypes = [vp, pd, pd, pd, C.POINTER(i), C.POINTER(C.c_int32)]
    for n in ("twin_multi_cur_time", "twin_multi_run_dur"):
        getattr(L, n).restype = d
        getattr(L, n).argtypes = [vp]
    L.twin_multi_ok.argtypes = [vp]
    return L


class TwinMulti(object):
    def __init__(self, S, capacity=1 << 15):
        self.L = _lib()
        ids = np.asarray(oracle.feature_ids(), dtype=np.int32)
        self.S = S
        self.h = self.L.twin_multi_create(S, 10, ids.ctypes.data_as(C.POINTER(C.c_int)), len(ids), capacity)

    def __del__(self):
        self.L.twin_multi_destroy(self.h)

    def reset(self, seed, bw, lat, queue, loss, rates):
        self.L.twin_multi_seed(self.h, int(seed))
        r = np.ascontiguousarray(rates, dtype=np.float64)
        self.L.twin_multi_reset(self.h, bw, lat, int(queue), loss, r.ctypes.data_as(C.POINTER(C.c_double)))

    def step(self, actions):
        a = np.ascontiguousarray(actions, dtype=np.float64)
        obs = np.zeros((self.S, 30)); rew = np.zeros(self.S); cnt = np.zeros((self.S, 3), dtype=np.int32)
        dn = C.c_int()
        pd = C.POINTER(C.c_double)
        self.L.twin_multi_step(self.h, a.ctypes.data_as(pd), obs.ctypes.data_as(pd), rew.ctypes.data_as(pd), C.byref(dn),
                               cnt.ctypes.data_as(C.POINTER(C.c_int32)))
        return obs, rew, bool(dn.value), cnt

    cur_time = property(lambda self: self.L.twin_multi_cur_time(self.h))
    run_dur = property(lambda self: self.L.twin_multi_run_dur(self.h))
    ok = property(lambda self: bool(self.L.twin_multi_ok(self.h)))


class TwinMultiFast(TwinMulti):
    """pcc_multi_fast.cuh: the same semantics without the event heap (shared ring, merged timers, three cursors)."""

    def __init__(self, S, capacity=1 << 15, ring_base=0):
        L = self.L = _lib()
        vp, d, i = C.c_void_p, C.c_double, C.c_int
        pd = C.POINTER(C.c_double)
        L.twin_mfast_create.restype = vp
        L.twin_mfast_create.argtypes = [i, i, C.POINTER(C.c_int), i, i, C.c_uint32]
        L.twin_mfast_destroy.argtypes = [vp]
        L.twin_mfast_seed.argtypes = [vp, C.c_uint64]
        L.twin_mfast_reset.argtypes = [vp, d, d, C.c_int64, d, pd]
        L.twin_mfast_step.argtypes = [vp, pd, pd, pd, C.POINTER(i), C.POINTER(C.c_int32)]
        for n in ("twin_mfast_cur_time", "twin_mfast_run_dur"):
            getattr(L, n).restype = d
            getattr(L, n).argtypes = [vp]
        L.twin_mfast_ok.argtypes = [vp]
        ids = np.asarray(oracle.feature_ids(), dtype=np.int32)
        self.S = S
        self.h = L.twin_mfast_create(S, 10, ids.ctypes.data_as(C.POINTER(C.c_int)), len(ids), capacity, ring_base)

    def __del__(self):
        self.L.twin_mfast_destroy(self.h)

    def reset(self, seed, bw, lat, queue, loss, rates):
        self.L.twin_mfast_seed(self.h, int(seed))
        r = np.ascontiguousarray(rates, dtype=np.float64)
        self.L.twin_mfast_reset(self.h, bw, lat, int(queue), loss, r.ctypes.data_as(C.POINTER(C.c_double)))

    def step(self, actions):
        a = np.ascontiguousarray(actions, dtype=np.float64)
        obs = np.zeros((self.S, 30)); rew = np.zeros(self.S); cnt = np.zeros((self.S, 3), dtype=np.int32)
        dn = C.c_int()
        pd = C.POINTER(C.c_double)
        self.L.twin_mfast_step(self.h, a.ctypes.data_as(pd), obs.ctypes.data_as(pd), rew.ctypes.data_as(pd), C.byref(dn),
                               cnt.ctypes.data_as(C.POINTER(C.c_int32)))
        return obs, rew, bool(dn.value), cnt

    cur_time = property(lambda self: self.L.twin_mfast_cur_time(self.h))
    run_dur = property(lambda self: self.L.twin_mfast_run_dur(self.h))
    ok = property(lambda self: bool(self.L.twin_mfast_ok(self.h)))


IMPLS = {"heap": TwinMulti, "stream": TwinMultiFast}


@pytest.mark.parametrize("impl", ["heap", "stream"])
@pytest.mark.parametrize("name", golden_names("multi_"))
def test_twin_multi_matches_reference_golden(name, impl):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    bw, lat, queue, loss = z["params"]
    t = IMPLS[impl](len(z["rates"]))
    t.reset(int(z["seed"]), bw, lat, int(queue), loss, z["rates"])
    assert t.cur_time == float(z["cur_time0"])
    for k in range(len(z["action"])):
        obs, rew, done, cnt = t.step(z["action"][k])
        assert np.array_equal(cnt, z["counts"][k]), (name, k)
        assert np.array_equal(obs, z["obs"][k]) and np.array_equal(rew, z["reward"][k]), (name, k)
        assert t.cur_time == z["cur_time"][k] and t.run_dur == z["run_dur"][k]
    assert t.ok


@pytest.mark.parametrize("impl", ["heap", "stream"])
@pytest.mark.parametrize("seed", range(8))
def test_twin_multi_equals_oracle_on_grid_points(seed, impl):
    g = np.random.default_rng(300 + seed)
    for trial in range(3):
        S = int(g.choice([2, 2, 3, 4]))
        bw = float(np.exp(g.uniform(np.log(83.0), np.log(83333.0))))      # 1..1000 Mbit/s
        lat = float(np.exp(g.uniform(np.log(0.001), np.log(0.5))))        # 1..500 ms
        queue = 1 + int(np.exp(g.uniform(0, 6)))
        loss = float(g.choice([0.0, 0.02, 0.3]))
        rates = g.uniform(40, 1000, S)
        o = oracle.OracleEnv()
        o.seed_philox(seed)
        o.reset_multi(bw, lat, queue, loss, rates)
        t = IMPLS[impl](S)
        t.reset(seed, bw, lat, queue, loss, rates)
        assert o.cur_time == t.cur_time
        for k in range(150):
            a = g.normal(0, 2.5, S)
            x = o.step_multi(a)
            y = t.step(a)
            assert np.array_equal(x[3], y[3]), (seed, trial, k)
            assert np.array_equal(x[0], y[0]) and np.array_equal(x[1], y[1]) and x[2] == y[2], (seed, trial, k)
            assert o.cur_time == t.cur_time and o.run_dur == t.run_dur
        assert t.ok


# ---- the cwnd / latency-noise variants of the loop on the same heap path ---------------
def _variant_lib():
    L = _lib()
    pd = C.POINTER(C.c_double)
    L.twin_multi_set_variant.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.twin_multi_cwnd.argtypes = [C.c_void_p, C.c_int]
    L.twin_multi_step_cwnd.argtypes = [C.c_void_p, pd, pd, pd, pd, C.POINTER(C.c_int), C.POINTER(C.c_int32)]
    return L


class TwinVariant(TwinMulti):
    def __init__(self, S, use_cwnd, use_noise, n_obs=30, features=None):
        self.L = _variant_lib()
        ids = np.asarray(oracle.feature_ids(features) if features else oracle.feature_ids(), dtype=np.int32)
        self.S, self.hf = S, 10 * len(ids)
        self.h = self.L.twin_multi_create(S, 10, ids.ctypes.data_as(C.POINTER(C.c_int)), len(ids), 1 << 15)
        self.L.twin_multi_set_variant(self.h, int(use_cwnd), int(use_noise))
        self.use_cwnd = use_cwnd

    def step2(self, actions, cwnd_actions):
        pd = C.POINTER(C.c_double)
        a = np.ascontiguousarray(actions, dtype=np.float64)
        ca = np.ascontiguousarray(cwnd_actions, dtype=np.float64)
        obs = np.zeros((self.S, self.hf)); rew = np.zeros(self.S); cnt = np.zeros((self.S, 3), dtype=np.int32)
        dn = C.c_int()
        self.L.twin_multi_step_cwnd(self.h, a.ctypes.data_as(pd), ca.ctypes.data_as(pd) if self.use_cwnd else None,
                                    obs.ctypes.data_as(pd), rew.ctypes.data_as(pd), C.byref(dn),
                                    cnt.ctypes.data_as(C.POINTER(C.c_int32)))
        return obs, rew, bool(dn.value), cnt

    def cwnd(self, i=0):
        return self.L.twin_multi_cwnd(self.h, i)


SINGLE_VARIANTS = [n for n in golden_names("variant_") if not n.startswith("variant_multi_")]


@pytest.mark.parametrize("name", SINGLE_VARIANTS)
def test_twin_variant_single_sender_golden(name):
    """USE_CWND / USE_LATENCY_NOISE on the reference's own single-sender env = this path with one sender."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    use_cwnd, use_noise = bool(z["use_cwnd"]), bool(z["use_noise"])
    n_steps = int(z["steps_per_episode"])
    k = 0
    for ep in range(len(z["ep_params"])):
        bw, lat, q, loss, rate = z["ep_params"][ep]
        t = TwinVariant(1, use_cwnd, use_noise, features=str(z["features"]))
        # one Philox stream runs through all episodes of the file: continue it by replaying the draw count
        if ep == 0:
            t.reset(int(z["seed"]), bw, lat, int(q), loss, [rate])
            keep = t
        else:
            t = keep
            t.L.twin_multi_reset(t.h, bw, lat, int(q), loss, np.array([rate]).ctypes.data_as(C.POINTER(C.c_double)))
        assert t.cur_time == z["ep_cur_time0"][ep]
        for _ in range(n_steps):
            obs, rew, done, cnt = t.step2([z["action"][k]], [z["cwnd_action"][k]])
            assert tuple(cnt[0]) == tuple(z["counts"][k]), (name, k)
            assert np.array_equal(obs[0], z["obs"][k]) and rew[0] == z["reward"][k], (name, k)
            assert t.cur_time == z["cur_time"][k] and t.run_dur == z["run_dur"][k]
            assert t.cwnd() == z["cwnd"][k]
            k += 1
        assert t.ok


@pytest.mark.parametrize("name", golden_names("variant_multi_"))
def test_twin_variant_multi_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    bw, lat, queue, loss = z["params"]
    S = len(z["rates"])
    t = TwinVariant(S, bool(z["use_cwnd"]), bool(z["use_noise"]))
    t.reset(int(z["seed"]), bw, lat, int(queue), loss, z["rates"])
    assert t.cur_time == float(z["cur_time0"])
    for k in range(len(z["action"])):
        obs, rew, done, cnt = t.step2(z["action"][k], z["cwnd_action"][k])
        assert np.array_equal(cnt, z["counts"][k]), (name, k)
        assert np.array_equal(obs, z["obs"][k]) and np.array_equal(rew, z["reward"][k]), (name, k)
        assert t.cur_time == z["cur_time"][k] and t.run_dur == z["run_dur"][k]
        assert [t.cwnd(i) for i in range(S)] == list(z["cwnd"][k])
    assert t.ok


@pytest.mark.parametrize("seed", range(6))
def test_twin_variant_equals_oracle_random(seed):
    g = np.random.default_rng(900 + seed)
    for trial in range(3):
        S = int(g.choice([1, 1, 2, 3]))
        use_cwnd, use_noise = bool(g.integers(0, 2)), bool(g.integers(0, 2))
        if not (use_cwnd or use_noise):
            use_cwnd = True
        bw = float(g.uniform(100, 2000)); lat = float(g.uniform(0.005, 0.4))
        queue = 1 + int(np.exp(g.uniform(0, 6))); loss = float(g.choice([0.0, 0.02, 0.2]))
        rates = g.uniform(40, 1000, S)
        o = oracle.OracleEnv()
        o.set_variant(use_cwnd, use_noise)
        o.seed_philox(seed)
        o.reset_multi(bw, lat, queue, loss, rates)
        t = TwinVariant(S, use_cwnd, use_noise)
        t.reset(seed, bw, lat, queue, loss, rates)
        assert o.cur_time == t.cur_time
        for k in range(120):
            a, ca = g.normal(0, 2.5, S), g.normal(0, 6.0, S)
            x = o.step_multi(a, ca)
            y = t.step2(a, ca)
            assert np.array_equal(x[3], y[3]), (seed, trial, k)
            assert np.array_equal(x[0], y[0]) and np.array_equal(x[1], y[1]) and x[2] == y[2], (seed, trial, k)
            assert o.cur_time == t.cur_time and o.run_dur == t.run_dur
            assert [o.cwnd(i) for i in range(S)] == [t.cwnd(i) for i in range(S)]
        assert t.ok


@pytest.mark.parametrize("seed", range(10))
def test_twin_multi_stream_nasty_ties_and_wrap(seed):
    """The heap-free path where it is most fragile: senders with IDENTICAL rates (their timers tie exactly, so the
    sender index decides every send order and many hop ties), heavy random loss and tiny queues (long drop clusters),
    ring cursors starting just below the u32 wrap."""
    g = np.random.default_rng(7000 + seed)
    for trial in range(3):
        S = int(g.choice([2, 3, 4]))
        bw = float(np.exp(g.uniform(np.log(40.0), np.log(5000.0))))
        lat = float(np.exp(g.uniform(np.log(0.001), np.log(0.3))))
        queue = 1 + int(np.exp(g.uniform(0, 4)))
        loss = float(g.choice([0.0, 0.05, 0.3, 0.9]))
        r0 = float(g.uniform(40, 1000))
        rates = np.full(S, r0) if trial != 1 else np.array([r0, r0 * 2, r0, r0 * 0.5][:S])
        o = oracle.OracleEnv()
        o.seed_philox(seed)
        o.reset_multi(bw, lat, queue, loss, rates)
        t = TwinMultiFast(S, ring_base=0xFFFFFF00)
        t.reset(seed, bw, lat, queue, loss, rates)
        assert o.cur_time == t.cur_time
        for k in range(200):
            a = np.full(S, float(g.normal(0, 2.0))) if k % 3 else g.normal(0, 2.5, S)   # equal actions keep the rates tied
            x = o.step_multi(a)
            y = t.step(a)
            assert np.array_equal(x[3], y[3]), (seed, trial, k)
            assert np.array_equal(x[0], y[0]) and np.array_equal(x[1], y[1]) and x[2] == y[2], (seed, trial, k)
            assert o.cur_time == t.cur_time and o.run_dur == t.run_dur
        assert t.ok


@pytest.mark.parametrize("S", [1, 3, 4])
def test_twin_multi_stream_ragged_links_other_sender_counts(S):
    """1, 3 and 4 senders on ragged links incl. queues of 0-2 packets and zero-packet MIs (the same generator family as
    tests/test_gpu_multi.py::test_cuda_multi_other_sender_counts_vs_oracle)."""
    n, steps = 48, 30
    g = np.random.default_rng(200 + S)
    for i in range(n):
        bw, lat = float(g.uniform(80, 2000)), float(np.exp(g.uniform(np.log(0.002), np.log(0.6))))
        queue, loss = int(g.integers(0, 40)) if i % 3 else int(g.integers(0, 3)), float(g.choice([0.0, 0.01, 0.05, 1.0]))
        rates = g.uniform(40, 1500, S)
        o = oracle.OracleEnv()
        o.seed_philox(300 + i)
        o.reset_multi(bw, lat, queue, loss, rates)
        tw = TwinMultiFast(S, capacity=1 << 14)
        tw.reset(300 + i, bw, lat, queue, loss, rates)
        for t in range(steps):
            a = g.normal(0, 2.0, S)
            x, y = o.step_multi(a), tw.step(a)
            assert np.array_equal(x[3], y[3]), (S, i, t)
            assert np.array_equal(x[0], y[0]) and np.array_equal(x[1], y[1]), (S, i, t)
        assert tw.ok
