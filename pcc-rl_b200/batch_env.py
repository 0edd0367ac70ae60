"""PccBatchEnv: N independent PCC-RL simulated-network environments stepped in lock step on one
B200 by libpcc_b200.so.  The batched counterpart of the reference's SimulatedNetworkEnv
(gym/network_sim.py:344-496): same constructor arguments (history_len, features), same
reset()/step() meaning per env, vector-env conventions for the batch (auto-reset; the obs of
a finished env is the first obs of its next episode).

All buffers are torch.cuda tensors; torch is only the allocator / stream provider -- the step
itself is one hand-written CUDA kernel reached through the C ABI (include/pcc_b200.h).
"""
import ctypes as C

import numpy as np

from . import _lib
from . import sender_obs
from .params import LinkRanges, sample_link_params, validate_link_params


def _require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("pcc_rl_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch


class PccBatchEnv(object):
    def __init__(self, n_envs, history_len=10, features=sender_obs.DEFAULT_FEATURES, device=None,
                 rng="philox", seed=0, ranges=None, ring_capacity=None, max_steps=None,
                 global_offset=0, n_global=None, auto_reset=True, want_info=False):
        torch = _require_cuda()
        self.torch = torch
        self.L = _lib.load()
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.n_envs = int(n_envs)
        self.history_len = int(history_len)
        self.features = sender_obs.feature_names(features)
        self.feature_ids = sender_obs.feature_ids(features)
        self.obs_dim = self.history_len * len(self.feature_ids)
        self.ranges = ranges or LinkRanges()
        self.seed_base = int(seed)
        self.global_offset = int(global_offset)
        self.n_global = int(n_global) if n_global is not None else self.global_offset + self.n_envs
        self.auto_reset = auto_reset
        self.want_info = want_info

        cfg = _lib.PccConfig()
        self.L.pcc_default_config(C.byref(cfg))
        cfg.device = self.device.index if self.device.index is not None else torch.cuda.current_device()
        cfg.n_envs = self.n_envs
        cfg.history_len = self.history_len
        cfg.n_features = len(self.feature_ids)
        for i, fid in enumerate(self.feature_ids):
            cfg.feature_ids[i] = fid
        cfg.rng_kind = {"philox": _lib.PCC_RNG_PHILOX, "mt19937": _lib.PCC_RNG_MT19937}[rng]
        if max_steps is not None:
            cfg.consts.max_steps = int(max_steps)
        if ring_capacity is None:
            ring_capacity = self.L.pcc_ring_capacity_for(cfg.consts.max_rate, self.ranges.bw[0], self.ranges.lat[1],
                                                         float(self.ranges.max_queue_packets()))
        cfg.ring_capacity = int(ring_capacity)
        self.cfg = cfg
        self.max_steps = cfg.consts.max_steps
        sb, rb = C.c_uint64(), C.c_uint64()
        _lib.check(self.L.pcc_workspace_bytes(C.byref(cfg), C.byref(sb), C.byref(rb)))
        with torch.cuda.device(self.device):
            free, _total = torch.cuda.mem_get_info()
            if sb.value + rb.value > free:
                raise RuntimeError(
                    "in-flight rings need %.1f GiB (n_envs=%d x ring_capacity=%d x 16 B) but only %.1f GiB are "
                    "free; narrow the LinkRanges or pass ring_capacity" %
                    (rb.value / 2**30, self.n_envs, cfg.ring_capacity, free / 2**30))
            self.state_ws = torch.empty(sb.value, dtype=torch.uint8, device=self.device)
            self.ring_ws = torch.empty(rb.value, dtype=torch.uint8, device=self.device)
            self.h = C.c_void_p()
            _lib.check(self.L.pcc_create(C.byref(self.h), C.byref(cfg), self.state_ws.data_ptr(),
                                         self.ring_ws.data_ptr()))
            f64 = dict(dtype=torch.float64, device=self.device)
            self.obs = torch.empty((self.n_envs, self.obs_dim), **f64)
            self.reward = torch.empty(self.n_envs, **f64)
            self.done = torch.empty(self.n_envs, dtype=torch.uint8, device=self.device)
            self.counts = torch.empty((self.n_envs, 3), dtype=torch.int32, device=self.device)
            self.info = torch.empty((self.n_envs, _lib.PCC_INFO_WIDTH), **f64) if want_info else None
            self._actions = torch.empty(self.n_envs, **f64)
        # host-side bookkeeping (deterministic, so no device->host sync is ever needed for it)
        self._steps = np.zeros(self.n_envs, dtype=np.int64)
        self._episode = np.zeros(self.n_envs, dtype=np.int64)
        self._global_ids = np.arange(self.global_offset, self.global_offset + self.n_envs, dtype=np.int64)
        self.params = None
        self.seed(self.seed_base)

    # -- plumbing ---------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def _dev(self, a, dtype):
        t = self.torch.as_tensor(np.ascontiguousarray(a), dtype=dtype)
        return t.to(self.device, non_blocking=False)

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.L.pcc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self):
        """Synchronises and raises if any env's in-flight ring overflowed."""
        _lib.check(self.L.pcc_check(self.h, self._stream()))

    @property
    def launches(self):
        return self.L.pcc_launch_count(self.h)

    def column(self, name):
        out = self.torch.empty(self.n_envs, dtype=self.torch.float64, device=self.device)
        _lib.check(self.L.pcc_get_column(self.h, name.encode(), out.data_ptr(), self._stream()))
        return out

    # -- gym-like surface ---------------------------------------------------------------------
    def seed(self, seed=None, seeds=None):
        """Per-env loss-draw streams.  Default: seed_base + GLOBAL env id (shard-invariant)."""
        if seeds is None:
            base = self.seed_base if seed is None else int(seed)
            self.seed_base = base
            seeds = (np.uint64(base & 0xFFFFFFFFFFFFFFFF) + self._global_ids.astype(np.uint64))
        seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
        assert seeds.shape == (self.n_envs,)
        t = self.torch.from_numpy(seeds.view(np.int64)).to(self.device)
        _lib.check(self.L.pcc_seed(self.h, t.data_ptr(), None, self._stream()))
        self.torch.cuda.current_stream(self.device).synchronize()
        return [self.seed_base]

    def reset(self, mask=None, params=None):
        """Starts a new episode for every env (or those in the boolean numpy `mask`).  `params` may
        give explicit link parameters: dict of arrays bw, lat, queue, loss, start_rate (length n_envs).
        Sampled parameters are prepared ahead of time (see _prefetch_start), so an episode boundary
        costs the reset kernel and five small uploads, not a host-side sampling pass."""
        torch = self.torch
        sel = np.ones(self.n_envs, dtype=bool) if mask is None else np.asarray(mask, dtype=bool)
        if mask is not None and sel.all():
            mask = None
        pinned = None
        if params is None:
            pinned = self._prefetch_take(sel)
            params = pinned[0] if pinned is not None else self._sample(sel)
        else:
            validate_link_params(*[np.asarray(params[k])[sel] for k in ("bw", "lat", "queue", "loss", "start_rate")])
        if self.params is None:
            self.params = {k: np.array(v, copy=True) for k, v in params.items()}
        else:
            for k in self.params:
                self.params[k][sel] = np.asarray(params[k])[sel]
        m = None if mask is None else self._dev(sel.astype(np.uint8), torch.uint8)
        if pinned is not None:      # page-locked staging: asynchronous uploads on the current stream
            bw, lat, q, loss, rate = (pinned[1][k].to(self.device, non_blocking=True)
                                      for k in ("bw", "lat", "queue", "loss", "start_rate"))
        else:
            bw = self._dev(params["bw"], torch.float64)
            lat = self._dev(params["lat"], torch.float64)
            q = self._dev(params["queue"], torch.int64)
            loss = self._dev(params["loss"], torch.float64)
            rate = self._dev(params["start_rate"], torch.float64)
        _lib.check(self.L.pcc_reset(self.h, m.data_ptr() if m is not None else None, bw.data_ptr(), lat.data_ptr(),
                                    q.data_ptr(), loss.data_ptr(), rate.data_ptr(), self.obs.data_ptr(),
                                    self._stream()))
        # the parameter tensors must outlive the asynchronous kernel
        self._keep = (m, bw, lat, q, loss, rate, pinned)
        self._steps[sel] = 0
        self._episode[sel] += 1
        if self.auto_reset:
            self._prefetch_start()
        return self.obs

    # -- parameters of the NEXT episode, sampled by a host thread while the current episode runs ------------------
    def _prefetch_start(self):
        import threading
        torch = self.torch
        if getattr(self, "_pin_sets", None) is None:
            # two alternating sets of page-locked staging buffers, allocated ONCE: the sampling thread must not make
            # CUDA calls (a cudaHostAlloc there waits for the whole queued stream to drain and stalls the boundary)
            mk = lambda dt: torch.empty(self.n_envs, dtype=dt).pin_memory()
            self._pin_sets = [{k: mk(torch.int64 if k == "queue" else torch.float64)
                               for k in ("bw", "lat", "queue", "loss", "start_rate")} for _ in range(2)]
            self._pin_next = 0
        pin = self._pin_sets[self._pin_next]
        self._pin_next ^= 1
        episodes = self._episode.copy()
        box = {}

        def work():
            p = self._sample_for(np.ones(self.n_envs, dtype=bool), episodes)
            for k, v in p.items():
                np.copyto(pin[k].numpy(), v)
            box["pin"] = pin
            box["host"] = p
        th = threading.Thread(target=work, daemon=True)
        th.start()
        self._pref = (th, episodes, box)

    def _prefetch_take(self, sel):
        pref, self._pref = getattr(self, "_pref", None), None
        if pref is None:
            return None
        th, episodes, box = pref
        th.join()
        if "host" not in box or not np.array_equal(episodes[sel], self._episode[sel]):
            return None
        return box["host"], box["pin"]

    def _sample(self, sel):
        """Reference-formula link parameters as a function of (seed, episode, global env id)."""
        return self._sample_for(sel, self._episode)

    def _sample_for(self, sel, episode):
        out = {k: np.zeros(self.n_envs, dtype=np.int64 if k == "queue" else np.float64)
               for k in ("bw", "lat", "queue", "loss", "start_rate")}
        idx = np.nonzero(sel)[0]
        for ep in np.unique(episode[idx]):
            ii = idx[episode[idx] == ep]
            p = sample_link_params(self.seed_base, int(ep), self._global_ids[ii], self.n_global, self.ranges)
            for k in out:
                out[k][ii] = p[k]
        return out

    def step(self, actions):
        """actions: tensor/array of shape [n_envs] or [n_envs, 1] (any float dtype; converted to
        float64 first, see SURVEY.md hard part 9).  Returns (obs[N, H*F], reward[N], done[N] bool, info)."""
        torch = self.torch
        a = actions
        if not torch.is_tensor(a):
            a = torch.as_tensor(np.asarray(a, dtype=np.float64))
        a = a.reshape(self.n_envs)
        if a.dtype == torch.float64 and a.device == self.device and a.is_contiguous():
            self._keep_actions = a            # already where the kernel reads it: no staging copy
        else:
            self._actions.copy_(a, non_blocking=True)
            a = self._actions
        _lib.check(self.L.pcc_step(self.h, a.data_ptr(), self.obs.data_ptr(), self.reward.data_ptr(),
                                   self.done.data_ptr(), self.counts.data_ptr(),
                                   self.info.data_ptr() if self.info is not None else None, self._stream()))
        self._steps += 1
        info = {"counts": self.counts}
        if self.info is not None:
            info["metrics"] = self.info
        finished = self._steps >= self.max_steps
        done = self.done.bool()
        if self.auto_reset and finished.any():
            # once per episode: surface ring overflows (the count on the device survives the reset).  AFTER the reset is
            # queued: the check waits for the stream, and the GPU should not sit idle while the host prepares the reset
            self.reset(mask=finished)
            self.check()
        return self.obs, self.reward, done, info

    def rollout(self, n_steps, actions=None, policy=None, want_obs=True, want_counts=True):
        """n_steps monitor intervals for every env without returning to the host: per step the policy / value
        kernel, the regular env step and the auto-reset of finished envs are enqueued back to back on the stream.
        Equivalent, bit for bit, to n_steps calls of step().

        actions: [n_steps, n_envs] float64 cuda tensor, or None to use `policy` = dict(w1, b1, w2, b2, w3,
        b3: float64 cuda tensors of the MLP obs -> h1 -> h2 -> 1 with tanh hidden layers; log_std, stochastic,
        noise_seed optional; vw1 ... vb3 optional: the value network of the same shape).  Returns dict(obs [K,N,H*F]
        (observation after each step), actions [K,N], reward [K,N], done [K,N] bool, counts [K,N,3], and with a value
        head vpred [K+1,N]: V(observation before step k), last row = V(observation after the last step))."""
        torch = self.torch
        K, n = int(n_steps), self.n_envs
        # parameters of the episodes that START during this rollout (host bookkeeping is deterministic)
        n_resets = (self._steps + K) // self.max_steps
        n_eps = int(n_resets.max()) if n > 0 else 0
        bank = np.zeros((max(n_eps, 1), 5, n))
        for j in range(n_eps):
            sel = n_resets > j
            pj = self._sample_for(sel, self._episode + j)
            for r, k in enumerate(("bw", "lat", "queue", "loss", "start_rate")):
                bank[j, r, sel] = pj[k][sel]
                self.params[k][sel] = pj[k][sel]
        bank_dev = torch.as_tensor(bank, dtype=torch.float64).to(self.device)
        f64 = dict(dtype=torch.float64, device=self.device)
        out = dict(reward=torch.empty((K, n), **f64), done=torch.empty((K, n), dtype=torch.uint8, device=self.device),
                   actions=torch.empty((K, n), **f64))
        out["obs"] = torch.empty((K, n, self.obs_dim), **f64) if want_obs else None
        out["counts"] = torch.empty((K, n, 3), dtype=torch.int32, device=self.device) if want_counts else None
        act_ptr, pol_ref, keep, vpred = None, None, [], None
        if actions is not None:
            a = torch.as_tensor(actions).to(self.device, torch.float64).reshape(K, n).contiguous()
            keep.append(a)
            act_ptr = a.data_ptr()
        if policy is not None:
            pol = _lib.PccPolicy()
            names = ["w1", "b1", "w2", "b2", "w3", "b3"]
            if "vw1" in policy:
                names += ["vw1", "vb1", "vw2", "vb2", "vw3", "vb3"]
                vpred = torch.empty((K + 1, n), **f64)
            ws = {k: policy[k].to(self.device, torch.float64).contiguous() for k in names}
            keep.append(ws)
            for k, v in ws.items():
                setattr(pol, k, v.data_ptr())
            pol.n_in, pol.h1, pol.h2 = ws["w1"].shape[1], ws["w1"].shape[0], ws["w2"].shape[0]
            assert ws["w2"].shape[1] == pol.h1 and ws["w3"].numel() == pol.h2 and pol.n_in == self.obs_dim
            if vpred is not None:
                assert ws["vw1"].shape == ws["w1"].shape and ws["vw2"].shape == ws["w2"].shape and ws["vw3"].numel() == pol.h2
            pol.log_std = float(policy.get("log_std", 0.0))
            pol.stochastic = int(bool(policy.get("stochastic", False)))
            pol.noise_seed = int(policy.get("noise_seed", 0)) & 0xFFFFFFFFFFFFFFFF
            pol_ref = C.byref(pol)
        p = lambda t: t.data_ptr() if t is not None else None
        _lib.check(self.L.pcc_rollout(self.h, K, act_ptr, pol_ref, bank_dev.data_ptr() if n_eps > 0 else None, n_eps,
                                      p(out["obs"]), out["actions"].data_ptr(), out["reward"].data_ptr(),
                                      out["done"].data_ptr(), p(out["counts"]), p(vpred), self._stream()))
        self._keep = (bank_dev, keep)
        self._steps = (self._steps + K) % self.max_steps
        self._episode += n_resets
        if want_obs:
            self.obs.copy_(out["obs"][K - 1])
        out["done"] = out["done"].bool()
        if vpred is not None:
            out["vpred"] = vpred
        return out

    def step_device(self, actions_f64):
        """Lowest-overhead step: `actions_f64` is a contiguous float64 cuda tensor [n_envs]; no
        auto-reset, no host bookkeeping beyond the step counter.  Returns None (read env.obs etc.)."""
        _lib.check(self.L.pcc_step(self.h, actions_f64.data_ptr(), self.obs.data_ptr(), self.reward.data_ptr(),
                                   self.done.data_ptr(), self.counts.data_ptr(),
                                   self.info.data_ptr() if self.info is not None else None, self._stream()))
        self._steps += 1

    def step_host(self, actions_np, obs_np, reward_np, done_np, counts_np=None):
        """The gym-style call with HOST buffers (numpy, ideally page-locked): actions go to the
        device, the MI runs, obs/reward/done come back; synchronous.  No auto-reset."""
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        _lib.check(self.L.pcc_step_host(self.h, p(actions_np), p(obs_np), p(reward_np), p(done_np),
                                        p(counts_np) if counts_np is not None else None, self._stream()))
        self._steps += 1

    def step_host_submit(self, actions_np, obs_np, reward_np, done_np, counts_np=None):
        """step_host split in two (pcc_step_host_submit / _wait): returns a ticket at once; the host buffers of a
        ticket are complete after step_host_wait(ticket).  Two tickets may be in flight, each with its own buffers."""
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        t = C.c_int64()
        _lib.check(self.L.pcc_step_host_submit(self.h, p(actions_np), p(obs_np), p(reward_np), p(done_np),
                                               p(counts_np) if counts_np is not None else None, None, self._stream(),
                                               C.byref(t)))
        self._steps += 1
        return t.value

    def step_host_wait(self, ticket):
        _lib.check(self.L.pcc_step_host_wait(self.h, int(ticket)))

    # -- observation / action space metadata (network_sim.py:376-388) -------------------------
    @property
    def single_observation_bounds(self):
        lo = np.tile(sender_obs.get_min_obs_vector(self.features), self.history_len)
        hi = np.tile(sender_obs.get_max_obs_vector(self.features), self.history_len)
        return lo, hi

    single_action_bounds = (np.array([-1e12]), np.array([1e12]))
