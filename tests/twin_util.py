"""ctypes wrapper over tests/twin/libpcc_twin.so: the product's pcc_core.cuh compiled for the
host (test harness; see tests/twin/pcc_twin.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

import oracle

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "twin", "pcc_twin.cpp")
_CSRC = os.path.join(os.path.dirname(_HERE), "pcc-rl_b200", "csrc")
_CORE = os.path.join(_CSRC, "pcc_core.cuh")
_DEPS = [os.path.join(_CSRC, f) for f in ("pcc_core.cuh", "pcc_multi_core.cuh", "pcc_multi_fast.cuh", "pcc_flows_core.cuh")]
_LIB = os.path.join(_HERE, "twin", "libpcc_twin.so")


def build(force=False):
    if (not force and os.path.exists(_LIB)
            and all(os.path.getmtime(_LIB) >= os.path.getmtime(s) for s in [_SRC] + _DEPS)):
        return _LIB
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wall",
                           "-Wno-unknown-pragmas", "-o", _LIB, _SRC])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        vp, d, i = C.c_void_p, C.c_double, C.c_int
        L.twin_create.restype = vp
        L.twin_create.argtypes = [i, C.POINTER(C.c_int), i, i]
        L.twin_destroy.argtypes = [vp]
        L.twin_seed_philox.argtypes = [vp, C.c_uint64]
        L.twin_mt_setstate.argtypes = [vp, C.POINTER(C.c_uint32)]
        L.twin_mt_getstate.argtypes = [vp, C.POINTER(C.c_uint32)]
        L.twin_set_max_steps.argtypes = [vp, i]
        L.twin_set_ring_cursor.argtypes = [vp, C.c_uint32]
        L.twin_set_prescan.argtypes = [vp, i]
        L.twin_get_obs.argtypes = [vp, C.POINTER(d)]
        L.twin_reset.argtypes = [vp, d, d, C.c_int64, d, d]
        L.twin_step.argtypes = [vp, d, C.POINTER(d), C.POINTER(d), C.POINTER(i),
                                C.POINTER(C.c_int64), C.POINTER(d)]
        for n in ("twin_cur_time", "twin_run_dur", "twin_rate"):
            getattr(L, n).restype = d
            getattr(L, n).argtypes = [vp]
        L.twin_overflow.argtypes = [vp]
        L.twin_inflight.restype = C.c_uint32
        L.twin_inflight.argtypes = [vp]
        _lib = L
    return _lib


class TwinEnv(object):
    def __init__(self, history_len=10, features=oracle.DEFAULT_FEATURES, ring_capacity=1 << 16):
        self.L = lib()
        ids = np.asarray(oracle.feature_ids(features), dtype=np.int32)
        self.h = self.L.twin_create(history_len, ids.ctypes.data_as(C.POINTER(C.c_int)), len(ids),
                                    ring_capacity)
        self._obs = np.zeros(history_len * len(ids))
        self._info = np.zeros(8)
        self._counts = np.zeros(3, dtype=np.int64)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.twin_destroy(self.h)
            self.h = None

    def seed_philox(self, seed):
        self.L.twin_seed_philox(self.h, int(seed) & 0xFFFFFFFFFFFFFFFF)

    def mt_setstate(self, st):
        st = np.ascontiguousarray(st, dtype=np.uint32)
        self.L.twin_mt_setstate(self.h, st.ctypes.data_as(C.POINTER(C.c_uint32)))

    def mt_getstate(self):
        st = np.zeros(625, dtype=np.uint32)
        self.L.twin_mt_getstate(self.h, st.ctypes.data_as(C.POINTER(C.c_uint32)))
        return st

    def set_prescan(self, on=True):
        """Two-stage consumption: the cursor scans once before the sends, then resumed (see pcc_core.cuh run_mi)."""
        self.L.twin_set_prescan(self.h, int(on))

    def set_ring_cursor(self, base):
        self.L.twin_set_ring_cursor(self.h, base)

    def reset(self, bw, lat, queue, loss, start_rate):
        self.L.twin_reset(self.h, bw, lat, int(queue), loss, start_rate)
        self.L.twin_get_obs(self.h, self._obs.ctypes.data_as(C.POINTER(C.c_double)))
        return self._obs.copy()

    def step(self, action):
        r, dn = C.c_double(), C.c_int()
        pd = C.POINTER(C.c_double)
        self.L.twin_step(self.h, float(action), self._obs.ctypes.data_as(pd), C.byref(r), C.byref(dn),
                         self._counts.ctypes.data_as(C.POINTER(C.c_int64)), self._info.ctypes.data_as(pd))
        return self._obs.copy(), r.value, bool(dn.value), self._counts.copy(), self._info.copy()

    cur_time = property(lambda self: self.L.twin_cur_time(self.h))
    run_dur = property(lambda self: self.L.twin_run_dur(self.h))
    rate = property(lambda self: self.L.twin_rate(self.h))
    overflow = property(lambda self: bool(self.L.twin_overflow(self.h)))
    inflight = property(lambda self: self.L.twin_inflight(self.h))


class TwinFlow(object):
    """One flow of the MI-sample ingestion path: pcc_flows_core.cuh compiled for the host."""

    def __init__(self, history_len=10, features=oracle.DEFAULT_FEATURES):
        L = self.L = lib()
        vp, d, i, q = C.c_void_p, C.c_double, C.c_int, C.c_int64
        L.twin_flow_create.restype = vp
        L.twin_flow_create.argtypes = [i, C.POINTER(C.c_int), i]
        L.twin_flow_destroy.argtypes = [vp]
        L.twin_flow_reset.argtypes = [vp, i]
        L.twin_flow_give_sample.argtypes = [vp, q, q, q, d, d, d, d, C.POINTER(d), q, q, C.POINTER(d)]
        L.twin_flow_get_obs.argtypes = [vp, C.POINTER(d)]
        L.twin_flow_apply_rate_delta.restype = d
        L.twin_flow_apply_rate_delta.argtypes = [d, d, d, d, d, i]
        ids = np.asarray(oracle.feature_ids(features), dtype=np.int32)
        self.hf = history_len * len(ids)
        self.h = L.twin_flow_create(history_len, ids.ctypes.data_as(C.POINTER(C.c_int)), len(ids))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.twin_flow_destroy(self.h)
            self.h = None

    def reset(self, mode):
        self.L.twin_flow_reset(self.h, mode)

    def give_sample(self, r):
        rtt = np.ascontiguousarray(r["rtt"], dtype=np.float64)
        m = np.zeros(12)
        self.L.twin_flow_give_sample(self.h, r["bytes_sent"], r["bytes_acked"], r["bytes_lost"], r["send_start"],
                                     r["send_end"], r["recv_start"], r["recv_end"],
                                     rtt.ctypes.data_as(C.POINTER(C.c_double)), rtt.size, r["packet_size"],
                                     m.ctypes.data_as(C.POINTER(C.c_double)))
        return m

    def obs(self):
        o = np.zeros(self.hf)
        self.L.twin_flow_get_obs(self.h, o.ctypes.data_as(C.POINTER(C.c_double)))
        return o

    def apply_rate_delta(self, rate, action, cfg):
        return self.L.twin_flow_apply_rate_delta(rate, action, cfg["delta_scale"], cfg["min_rate"], cfg["max_rate"],
                                                 cfg["style"])
