"""PccMultiSenderEnv: N independent links, each shared by S senders (BASELINE config 5: the bw x delay grid
sweep with 2 senders per link).  Batched counterpart of driving the reference's Network with several Sender
objects (gym/network_sim.py:100-178); semantics in include/pcc_b200.h / DESIGN.md.  Default engine: the heap-free
streaming MI with one link per warp (csrc/pcc_multi_warp.cuh); the per-env event heap serves the variants below.

The same path carries the two variants the reference hides behind module switches (network_sim.py:51-54):
`use_cwnd=True` (USE_CWND: congestion window, a second action component per sender) and
`use_latency_noise=True` (USE_LATENCY_NOISE: per-hop latency jitter).  With n_senders=1 this is the reference's
own SimulatedNetworkEnv with the switch turned on (tests/golden/variant_*.npz)."""
import ctypes as C

import numpy as np

from . import _lib, sender_obs
from .params import validate_link_params


def grid_sweep_params(bw_mbps=(1.0, 1000.0), lat_ms=(1.0, 500.0), n_bw=32, n_lat=32, queue=50, loss=0.0,
                      bytes_per_packet=1500):
    """Log-spaced grid of the sweep; bandwidth in the reference's unit (packets/s of 1500 B)."""
    bw = np.exp(np.linspace(np.log(bw_mbps[0]), np.log(bw_mbps[1]), n_bw)) * 1e6 / (8 * bytes_per_packet)
    lat = np.exp(np.linspace(np.log(lat_ms[0]), np.log(lat_ms[1]), n_lat)) * 1e-3
    B, Lm = np.meshgrid(bw, lat, indexing="ij")
    n = B.size
    return dict(bw=B.reshape(-1), lat=Lm.reshape(-1), queue=np.full(n, int(queue), dtype=np.int64),
                loss=np.full(n, float(loss)))


class PccMultiSenderEnv(object):
    def __init__(self, n_envs, n_senders=2, history_len=10, features=sender_obs.DEFAULT_FEATURES, device=None, seed=0,
                 ring_capacity=8192, max_steps=None, use_cwnd=False, use_latency_noise=False):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("pcc_rl_b200 needs a CUDA device; there is no CPU fallback")
        self.torch, self.L = torch, _lib.load()
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.n_envs, self.S = int(n_envs), int(n_senders)
        self.feature_ids = sender_obs.feature_ids(features)
        self.obs_dim = history_len * len(self.feature_ids)
        cfg = _lib.PccConfig()
        self.L.pcc_default_config(C.byref(cfg))
        cfg.device = self.device.index if self.device.index is not None else torch.cuda.current_device()
        cfg.n_envs, cfg.history_len, cfg.n_features = self.n_envs, history_len, len(self.feature_ids)
        for i, fid in enumerate(self.feature_ids):
            cfg.feature_ids[i] = fid
        cfg.ring_capacity = int(ring_capacity)
        if max_steps is not None:
            cfg.consts.max_steps = int(max_steps)
        self.cfg = cfg
        nb = C.c_uint64()
        _lib.check(self.L.pcc_multi_workspace_bytes(C.byref(cfg), self.S, C.byref(nb)))
        with torch.cuda.device(self.device):
            self.ws = torch.empty(nb.value, dtype=torch.uint8, device=self.device)
            self.h = C.c_void_p()
            _lib.check(self.L.pcc_multi_create(C.byref(self.h), C.byref(cfg), self.S, self.ws.data_ptr()))
            self.use_cwnd, self.use_latency_noise = bool(use_cwnd), bool(use_latency_noise)
            # the event-heap engine reports every sender's window (Sender.cwnd, the initial 25 when USE_CWND is off); the
            # streaming engines have none: `cwnd` stays zero and the per-step memset is skipped
            import os
            self._heap_engine = self.use_cwnd or self.use_latency_noise or os.environ.get("PCC_MULTI_MODE") == "heap"
            if use_cwnd or use_latency_noise:
                v = _lib.PccVariant()
                self.L.pcc_default_variant(C.byref(v))
                v.use_cwnd, v.use_latency_noise = int(self.use_cwnd), int(self.use_latency_noise)
                _lib.check(self.L.pcc_multi_set_variant(self.h, C.byref(v)))
            self.cwnd = torch.zeros((self.n_envs, self.S), dtype=torch.int32, device=self.device)
            f64 = dict(dtype=torch.float64, device=self.device)
            self.obs = torch.empty((self.n_envs, self.S, self.obs_dim), **f64)
            self.reward = torch.empty((self.n_envs, self.S), **f64)
            self.done = torch.empty(self.n_envs, dtype=torch.uint8, device=self.device)
            self.counts = torch.empty((self.n_envs, self.S, 3), dtype=torch.int32, device=self.device)
        self.seed(seed)

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def close(self):
        if getattr(self, "h", None):
            self.L.pcc_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def seed(self, seed=0, seeds=None):
        if seeds is None:
            seeds = np.uint64(int(seed) & 0xFFFFFFFFFFFFFFFF) + np.arange(self.n_envs, dtype=np.uint64)
        t = self.torch.from_numpy(np.ascontiguousarray(seeds, dtype=np.uint64).view(np.int64)).to(self.device)
        _lib.check(self.L.pcc_multi_seed(self.h, t.data_ptr(), self._stream()))
        self.torch.cuda.current_stream(self.device).synchronize()

    def reset(self, params, start_rates):
        """params: dict bw, lat, queue, loss (length n_envs); start_rates: [n_envs, n_senders] packets/s."""
        torch = self.torch
        validate_link_params(params["bw"], params["lat"], params["queue"], params["loss"], start_rates)
        dev = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).to(self.device)
        bw, lat, loss = (dev(params[k], torch.float64) for k in ("bw", "lat", "loss"))
        q = dev(params["queue"], torch.int64)
        r = dev(np.asarray(start_rates, dtype=np.float64).reshape(self.n_envs, self.S), torch.float64)
        _lib.check(self.L.pcc_multi_reset(self.h, None, bw.data_ptr(), lat.data_ptr(), q.data_ptr(), loss.data_ptr(),
                                          r.data_ptr(), self.obs.data_ptr(), self._stream()))
        self._keep = (bw, lat, loss, q, r)
        return self.obs

    def step(self, actions, cwnd_actions=None):
        """actions [n_envs, n_senders] (rate); with use_cwnd also cwnd_actions [n_envs, n_senders], or actions of
        shape [n_envs, n_senders, 2] = the reference's 2-dim action (network_sim.py:380-381, 412-414)."""
        torch = self.torch
        a = torch.as_tensor(actions).to(self.device, torch.float64)
        if self.use_cwnd and cwnd_actions is None and a.dim() >= 2 and a.shape[-1] == 2 and a.numel() == 2 * self.n_envs * self.S:
            a, cwnd_actions = a.reshape(self.n_envs, self.S, 2)[..., 0], a.reshape(self.n_envs, self.S, 2)[..., 1]
        a = a.reshape(self.n_envs, self.S).contiguous()
        ca = None
        if self.use_cwnd:
            if cwnd_actions is None:
                raise ValueError("use_cwnd: step needs the window actions too")
            ca = torch.as_tensor(cwnd_actions).to(self.device, torch.float64).reshape(self.n_envs, self.S).contiguous()
        _lib.check(self.L.pcc_multi_step_cwnd(self.h, a.data_ptr(), ca.data_ptr() if ca is not None else None,
                                              self.obs.data_ptr(), self.reward.data_ptr(), self.done.data_ptr(),
                                              self.counts.data_ptr(), self.cwnd.data_ptr() if self._heap_engine else None,
                                              self._stream()))
        self._keep_a = (a, ca)
        return self.obs, self.reward, self.done.bool(), {"counts": self.counts, "cwnd": self.cwnd}

    def check(self):
        _lib.check(self.L.pcc_multi_check(self.h, self._stream()))

    def launch_count(self):
        return int(self.L.pcc_multi_launch_count(self.h))
