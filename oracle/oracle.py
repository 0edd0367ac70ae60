"""ctypes wrapper over oracle/libpcc_oracle.so -- TEST INFRASTRUCTURE, not product code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (pcc-rl_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpcc_oracle.so")

# metric ids = position in the reference's SENDER_MI_METRICS (sender_obs.py:193-206)
METRIC_NAMES = ["send rate", "recv rate", "recv dur", "send dur", "avg latency", "loss ratio",
                "ack latency inflation", "sent latency inflation", "conn min latency",
                "latency increase", "latency ratio", "send ratio"]
DEFAULT_FEATURES = "sent latency inflation,latency ratio,send ratio"


def feature_ids(features=DEFAULT_FEATURES):
    names = features.split(",") if isinstance(features, str) else list(features)
    return [METRIC_NAMES.index(n) for n in names]


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("pcc_oracle.c", "pcc_oracle_batch.c", "pcc_oracle_flows.c", "Makefile")]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in src)):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B", "libpcc_oracle.so"],
                          stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        build()
    L = C.CDLL(_LIB_PATH)
    vp, d, i, l = C.c_void_p, C.c_double, C.c_int, C.c_long
    pd, pi, pl = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_long)
    L.pcco_create.restype = vp
    L.pcco_create.argtypes = [i, pi, i]
    L.pcco_destroy.argtypes = [vp]
    L.pcco_set_max_steps.argtypes = [vp, l]
    L.pcco_seed_mt.argtypes = [vp, C.c_uint64]
    L.pcco_seed_philox.argtypes = [vp, C.c_uint64]
    L.pcco_mt_getstate.argtypes = [vp, C.POINTER(C.c_uint32)]
    L.pcco_mt_setstate.argtypes = [vp, C.POINTER(C.c_uint32)]
    L.pcco_philox_draws.restype = C.c_uint64
    L.pcco_philox_draws.argtypes = [vp]
    L.pcco_random.restype = d
    L.pcco_random.argtypes = [vp]
    L.pcco_reset.argtypes = [vp, d, d, l, d, d]
    L.pcco_get_obs.argtypes = [vp, pd]
    L.pcco_step.argtypes = [vp, d, pd, pd, pi, pl, pd]
    for name in ("pcco_cur_time", "pcco_run_dur", "pcco_rate"):
        getattr(L, name).restype = d
        getattr(L, name).argtypes = [vp]
    L.pcco_queue_len.restype = l
    L.pcco_queue_len.argtypes = [vp]
    L.pcco_total_events.restype = C.c_longlong
    L.pcco_total_events.argtypes = [vp]
    L.pcco_reset_multi.argtypes = [vp, i, d, d, l, d, pd]
    L.pcco_step_multi.argtypes = [vp, pd, pd, pd, pi, pl]
    L.pcco_set_variant.argtypes = [vp, i, i]
    L.pcco_cwnd.restype = l
    L.pcco_cwnd.argtypes = [vp, i]
    L.pcco_step_cwnd.argtypes = [vp, d, d, pd, pd, pi, pl, pd]
    L.pcco_step_multi_cwnd.argtypes = [vp, pd, pd, pd, pd, pi, pl]
    L.pcco_np_mean.restype = d
    L.pcco_np_mean.argtypes = [pd, l]
    L.pcco_batch_run.restype = d
    L.pcco_batch_run.argtypes = [l, l, i, i, pi, i, pd, pd, pl, pd, pd,
                                 C.POINTER(C.c_uint64), pd, pd, pd, pl, pd, pi]
    _lib = L
    return L


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


def np_mean(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return lib().pcco_np_mean(_p(a, C.c_double), a.size)


class OracleEnv(object):
    """One env of the restated reference (1 sender, 2 links, network_sim.py:344-496)."""

    def __init__(self, history_len=10, features=DEFAULT_FEATURES):
        self.L = lib()
        ids = np.asarray(feature_ids(features), dtype=np.int32)
        self.history_len, self.n_features = history_len, len(ids)
        self.h = self.L.pcco_create(history_len, _p(ids, C.c_int), len(ids))
        if not self.h:
            raise ValueError("bad history_len / features")
        self._obs = np.zeros(history_len * len(ids))
        self._info = np.zeros(8)
        self._counts = np.zeros(3, dtype=np.int64)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.pcco_destroy(self.h)
            self.h = None

    def seed_mt(self, seed):
        self.L.pcco_seed_mt(self.h, int(seed))

    def seed_philox(self, seed):
        self.L.pcco_seed_philox(self.h, int(seed) & 0xFFFFFFFFFFFFFFFF)

    def mt_getstate(self):
        s = np.zeros(625, dtype=np.uint32)
        self.L.pcco_mt_getstate(self.h, _p(s, C.c_uint32))
        return s

    def mt_setstate(self, s):
        s = np.ascontiguousarray(s, dtype=np.uint32)
        assert s.size == 625
        self.L.pcco_mt_setstate(self.h, _p(s, C.c_uint32))

    def random(self):
        return self.L.pcco_random(self.h)

    def uniform(self, a, b):
        return a + (b - a) * self.random()

    def set_max_steps(self, n):
        self.L.pcco_set_max_steps(self.h, n)

    def set_variant(self, use_cwnd=False, use_latency_noise=False):
        """The reference's module switches USE_CWND / USE_LATENCY_NOISE (network_sim.py:51-54)."""
        self.L.pcco_set_variant(self.h, int(use_cwnd), int(use_latency_noise))

    def cwnd(self, sender=0):
        return self.L.pcco_cwnd(self.h, sender)

    def step_cwnd(self, action, cwnd_action):
        r, dn = C.c_double(), C.c_int()
        self.L.pcco_step_cwnd(self.h, float(action), float(cwnd_action), _p(self._obs, C.c_double), C.byref(r),
                              C.byref(dn), _p(self._counts, C.c_long), _p(self._info, C.c_double))
        return self._obs.copy(), r.value, bool(dn.value), self._counts.copy(), self._info.copy()

    def sample_params(self, bw=(100, 500), lat=(0.05, 0.5), queue=(0, 8), loss=(0.0, 0.05)):
        """The five draws of create_new_links_and_senders (network_sim.py:455-466), taken from
        this env's own stream, in the reference's order."""
        b = self.uniform(*bw)
        l = self.uniform(*lat)
        q = 1 + int(np.exp(self.uniform(*queue)))
        lo = self.uniform(*loss)
        r = self.uniform(0.3, 1.5) * b
        return b, l, q, lo, r

    def reset(self, bw, lat, queue, loss, start_rate):
        self.L.pcco_reset(self.h, bw, lat, int(queue), loss, start_rate)
        self.L.pcco_get_obs(self.h, _p(self._obs, C.c_double))
        return self._obs.copy()

    def step(self, action):
        r, dn = C.c_double(), C.c_int()
        self.L.pcco_step(self.h, float(action), _p(self._obs, C.c_double), C.byref(r), C.byref(dn),
                         _p(self._counts, C.c_long), _p(self._info, C.c_double))
        return self._obs.copy(), r.value, bool(dn.value), self._counts.copy(), self._info.copy()

    # -- several senders on one bottleneck (config 5) --
    def reset_multi(self, bw, lat, queue, loss, start_rates):
        r = np.ascontiguousarray(start_rates, dtype=np.float64)
        self.n_senders = len(r)
        self.L.pcco_reset_multi(self.h, len(r), bw, lat, int(queue), loss, _p(r, C.c_double))

    def step_multi(self, actions, cwnd_actions=None):
        S = self.n_senders
        a = np.ascontiguousarray(actions, dtype=np.float64)
        if cwnd_actions is not None:
            ca = np.ascontiguousarray(cwnd_actions, dtype=np.float64)
            obs = np.zeros((S, self.history_len * self.n_features))
            rew = np.zeros(S)
            cnt = np.zeros((S, 3), dtype=np.int64)
            dn = C.c_int()
            self.L.pcco_step_multi_cwnd(self.h, _p(a, C.c_double), _p(ca, C.c_double), _p(obs, C.c_double),
                                        _p(rew, C.c_double), C.byref(dn), _p(cnt, C.c_long))
            return obs, rew, bool(dn.value), cnt
        obs = np.zeros((S, self.history_len * self.n_features))
        rew = np.zeros(S)
        cnt = np.zeros((S, 3), dtype=np.int64)
        dn = C.c_int()
        self.L.pcco_step_multi(self.h, _p(a, C.c_double), _p(obs, C.c_double), _p(rew, C.c_double), C.byref(dn),
                               _p(cnt, C.c_long))
        return obs, rew, bool(dn.value), cnt

    @property
    def cur_time(self):
        return self.L.pcco_cur_time(self.h)

    @property
    def run_dur(self):
        return self.L.pcco_run_dur(self.h)

    @property
    def rate(self):
        return self.L.pcco_rate(self.h)

    @property
    def total_events(self):
        return self.L.pcco_total_events(self.h)


def batch_run(bw, lat, queue, loss, start_rate, seeds, n_steps, actions=None, n_threads=1,
              history_len=10, features=DEFAULT_FEATURES, trajectories=False):
    """Reset + n_steps steps for every env on n_threads host threads (Philox streams).
    Returns dict(seconds, obs[N,H*F] of the last step, reward_sum[N], count_sum[N,3])."""
    L = lib()
    n = len(bw)
    ids = np.asarray(feature_ids(features), dtype=np.int32)
    f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    bw, lat, loss, start_rate = f64(bw), f64(lat), f64(loss), f64(start_rate)
    queue = np.ascontiguousarray(queue, dtype=np.int64)
    seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
    obs = np.zeros((n, history_len * len(ids)))
    rs = np.zeros(n)
    cs = np.zeros((n, 3), dtype=np.int64)
    if actions is not None:
        actions = f64(actions)
        assert actions.shape == (n_steps, n)
    rt = np.zeros((n_steps, n)) if trajectories else None
    ct = np.zeros((n_steps, n, 3), dtype=np.int32) if trajectories else None
    secs = L.pcco_batch_run(n, n_steps, n_threads, history_len, _p(ids, C.c_int), len(ids),
                            _p(bw, C.c_double), _p(lat, C.c_double), _p(queue, C.c_long),
                            _p(loss, C.c_double), _p(start_rate, C.c_double),
                            _p(seeds, C.c_uint64),
                            _p(actions, C.c_double) if actions is not None else None,
                            _p(obs, C.c_double), _p(rs, C.c_double), _p(cs, C.c_long),
                            _p(rt, C.c_double) if trajectories else None,
                            _p(ct, C.c_int) if trajectories else None)
    return dict(seconds=secs, obs=obs, reward_sum=rs, count_sum=cs, reward_traj=rt, count_traj=ct)


class OracleBatch(object):
    """Persistent batch of oracle envs on host threads (the CPU arm of bench.py)."""

    def __init__(self, seeds, history_len=10, features=DEFAULT_FEATURES, n_threads=1):
        L = self.L = lib()
        L.pcco_batch_create.restype = C.c_void_p
        L.pcco_batch_create.argtypes = [C.c_long, C.c_int, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_uint64)]
        L.pcco_batch_destroy.argtypes = [C.c_void_p]
        vp = C.c_void_p
        L.pcco_batch_reset.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, C.c_int]
        L.pcco_batch_step.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int]
        seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
        ids = np.asarray(feature_ids(features), dtype=np.int32)
        self.n, self.hf, self.n_threads = len(seeds), history_len * len(ids), n_threads
        self.b = L.pcco_batch_create(self.n, history_len, _p(ids, C.c_int), len(ids), _p(seeds, C.c_uint64))
        self.obs = np.zeros((self.n, self.hf))
        self.reward = np.zeros(self.n)
        self.done = np.zeros(self.n, dtype=np.uint8)
        self.counts = np.zeros((self.n, 3), dtype=np.int32)

    def __del__(self):
        if getattr(self, "b", None):
            self.L.pcco_batch_destroy(self.b)
            self.b = None

    def reset(self, params, mask=None):
        v = lambda a, dt: np.ascontiguousarray(a, dtype=dt)
        bw, lat, loss, sr = (v(params[k], np.float64) for k in ("bw", "lat", "loss", "start_rate"))
        q = v(params["queue"], np.int64)
        m = None if mask is None else v(mask, np.uint8)
        self.L.pcco_batch_reset(self.b, m.ctypes.data if m is not None else None, bw.ctypes.data, lat.ctypes.data,
                                q.ctypes.data, loss.ctypes.data, sr.ctypes.data, self.obs.ctypes.data,
                                self.n_threads)
        return self.obs

    def step(self, actions):
        a = np.ascontiguousarray(actions, dtype=np.float64)
        self.L.pcco_batch_step(self.b, a.ctypes.data, self.obs.ctypes.data, self.reward.ctypes.data,
                               self.done.ctypes.data, self.counts.ctypes.data, self.n_threads)
        return self.obs, self.reward, self.done, self.counts


class OracleFlows(object):
    """The MI-sample ingestion path restated (oracle/pcc_oracle_flows.c): per-flow SenderHistory fed by
    give_sample records (loaded_client.py:111-138, shim_env.py:102-139)."""

    def __init__(self, n_flows, history_len=10, features=DEFAULT_FEATURES):
        L = self.L = lib()
        vp, d, i, l, ll = C.c_void_p, C.c_double, C.c_int, C.c_long, C.c_longlong
        pd = C.POINTER(C.c_double)
        L.pcco_flows_create.restype = vp
        L.pcco_flows_create.argtypes = [l, i, C.POINTER(C.c_int), i]
        L.pcco_flows_destroy.argtypes = [vp]
        L.pcco_flows_reset.argtypes = [vp, l, i]
        L.pcco_flows_set_rate.argtypes = [vp, l, d]
        L.pcco_flows_rate.restype = d
        L.pcco_flows_rate.argtypes = [vp, l]
        L.pcco_flows_conn_min.restype = d
        L.pcco_flows_conn_min.argtypes = [vp, l]
        L.pcco_flows_n_records.restype = ll
        L.pcco_flows_n_records.argtypes = [vp, l]
        L.pcco_flows_give_sample.argtypes = [vp, l, ll, ll, ll, d, d, d, d, pd, l, ll, pd]
        L.pcco_flows_get_obs.argtypes = [vp, l, pd]
        L.pcco_flows_apply_rate_delta.restype = d
        L.pcco_flows_apply_rate_delta.argtypes = [d, d, d, d, d, i]
        ids = np.asarray(feature_ids(features), dtype=np.int32)
        self.n_flows, self.history_len, self.n_features = n_flows, history_len, len(ids)
        self.h = L.pcco_flows_create(n_flows, history_len, _p(ids, C.c_int), len(ids))
        if not self.h:
            raise ValueError("bad n_flows / history_len / features")

    def __del__(self):
        if getattr(self, "h", None):
            self.L.pcco_flows_destroy(self.h)
            self.h = None

    def reset(self, flow, mode=0):
        self.L.pcco_flows_reset(self.h, flow, mode)

    def give_sample(self, flow, bytes_sent, bytes_acked, bytes_lost, send_start, send_end, recv_start, recv_end,
                    rtt_samples, packet_size, want_metrics=False):
        rtt = np.ascontiguousarray(rtt_samples, dtype=np.float64)
        m = np.zeros(12) if want_metrics else None
        self.L.pcco_flows_give_sample(self.h, flow, int(bytes_sent), int(bytes_acked), int(bytes_lost),
                                      float(send_start), float(send_end), float(recv_start), float(recv_end),
                                      _p(rtt, C.c_double), rtt.size, int(packet_size),
                                      _p(m, C.c_double) if want_metrics else None)
        return m

    def give_batch(self, b, n_threads=1):
        """b: dict in the layout of tests/flows_util.synth_batch; returns seconds spent in C."""
        import time
        v = lambda k, dt: np.ascontiguousarray(b[k], dtype=dt)
        a = [v("flow", np.int32)] + [v(k, np.int64) for k in ("bytes_sent", "bytes_acked", "bytes_lost")] + \
            [v(k, np.float64) for k in ("send_start", "send_end", "recv_start", "recv_end")] + \
            [v("packet_size", np.int64), v("rtt_off", np.int64), v("rtt", np.float64)]
        self.L.pcco_flows_give_batch.argtypes = [C.c_void_p, C.c_long] + [C.c_void_p] * 11
        self.L.pcco_flows_give_batch_mt.argtypes = [C.c_void_p, C.c_int, C.c_long] + [C.c_void_p] * 11
        t0 = time.perf_counter()
        if n_threads > 1:
            self.L.pcco_flows_give_batch_mt(self.h, int(n_threads), len(a[0]), *[x.ctypes.data for x in a])
        else:
            self.L.pcco_flows_give_batch(self.h, len(a[0]), *[x.ctypes.data for x in a])
        return time.perf_counter() - t0

    def obs(self, flow):
        o = np.zeros(self.history_len * self.n_features)
        self.L.pcco_flows_get_obs(self.h, flow, _p(o, C.c_double))
        return o

    def conn_min(self, flow):
        return self.L.pcco_flows_conn_min(self.h, flow)

    def n_records(self, flow):
        return self.L.pcco_flows_n_records(self.h, flow)

    def apply_rate_delta(self, rate, delta, delta_scale, min_rate, max_rate, style=0):
        return self.L.pcco_flows_apply_rate_delta(rate, delta, delta_scale, min_rate, max_rate, style)
