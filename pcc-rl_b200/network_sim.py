"""Drop-in for the reference's gym/network_sim.py module: `SimulatedNetworkEnv` with the same
constructor, reset/step/seed/render/close surface, observation_space/action_space and the
'PccNs-v0' registration (gym/network_sim.py:344-498) -- but every monitor interval runs on the
GPU through libpcc_b200.so (one env = a batch of 1; use PccBatchEnv for throughput).

Put this directory first on sys.path and stable_solve.py's `import network_sim` /
`gym.make('PccNs-v0')` pick this class up unchanged.

Fidelity: like the reference, all randomness comes from Python's global `random` module
(random.seed(s) seeds the env).  Link parameters are drawn on the host with the reference's
five random.uniform calls (:455-466); for the per-packet loss draws (:73) the global MT19937
state is handed to the device before a reset/step and read back afterwards, so the stream
continues exactly as in the reference.  Same seed -> bit-identical trajectories
(tests/test_gpu_dropin.py replays the reference's own outputs).
"""
import ctypes as C
import json
import random

import numpy as np

try:  # package import or top-level `import network_sim` (sys.path drop-in, tests/test_gpu_dropin.py)
    from . import _lib, sender_obs
except ImportError:
    import importlib
    import os
    import sys
    _here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.dirname(_here))
    _pkg = importlib.import_module("pcc_rl_b200")
    _lib, sender_obs = _pkg._lib, _pkg.sender_obs

try:
    import gym
    from gym import spaces
    from gym.utils import seeding
    from gym.envs.registration import register
    _EnvBase = gym.Env
except Exception:  # gym is optional: the env works without it
    gym = None
    _EnvBase = object

    class _Box(object):
        def __init__(self, low, high, dtype=np.float32):
            self.low = np.asarray(low, dtype=dtype)
            self.high = np.asarray(high, dtype=dtype)
            self.dtype = np.dtype(dtype)
            self.shape = self.low.shape

    class spaces(object):
        Box = _Box

    class seeding(object):
        @staticmethod
        def np_random(seed=None):
            return np.random.RandomState(seed), seed

def arg_or_default(arg, default=None):
    """The reference reads its constructor defaults from the command line (`stable_solve.py --history-len=5
    --input-features=...`, common/simple_arg_parse.py:17-35 -> network_sim.py:347-351): `name=value` tokens of sys.argv,
    converted to the type of the default.  Same rule here, so the flags keep working when this module is dropped in."""
    import sys
    for a in sys.argv:
        eq = a.find("=")
        key, val = (a[:eq], a[eq + 1:]) if eq >= 0 else (a, True)
        if key == arg:
            if isinstance(default, int):
                return int(val)
            if isinstance(default, float):
                return float(val)
            return val
    return default


# reference constants (gym/network_sim.py:33-54)
MAX_RATE = 1000
MIN_RATE = 40
REWARD_SCALE = 0.001
MAX_STEPS = 400
BYTES_PER_PACKET = 1500
# The reference's module switches (network_sim.py:51-54).  Shipped configuration (both False): the three-cursor kernels
# with Python's own MT19937 stream, bit-identical to the reference for the same random.seed().  With a switch set (here,
# before constructing the env, exactly as one would edit the reference's constants) the env runs on the per-env event-heap
# engine (PccMultiSenderEnv, one sender): the action space becomes 2-dimensional with USE_CWND (:376-379, 413-414) and
# the noise / loss draws come from a Philox stream seeded from Python's `random` at every reset (reproducible for a
# given random.seed(), pinned against the reference by tests/golden/variant_*.npz -- not the reference's own stream).
USE_CWND = False
USE_LATENCY_NOISE = False


class SimulatedNetworkEnv(_EnvBase):

    def __init__(self, history_len=arg_or_default("--history-len", default=10),
                 features=arg_or_default("--input-features",
                                         default="sent latency inflation,latency ratio,send ratio"),
                 device=None, strict_rng=False):
        """strict_rng=False: the MT19937 state stays on the device between steps and is exchanged with Python's global
        `random` only where the env itself draws from it (reset); strict_rng=True exchanges it around every step, for
        callers that draw from `random` between steps and need the reference's exact interleaving."""
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("pcc_rl_b200.SimulatedNetworkEnv needs a CUDA device; there is no CPU fallback")
        self.torch = torch
        self.viewer = None
        self.rand = None
        self.min_bw, self.max_bw = (100, 500)
        self.min_lat, self.max_lat = (0.05, 0.5)
        self.min_queue, self.max_queue = (0, 8)
        self.min_loss, self.max_loss = (0.0, 0.05)
        self.history_len = history_len
        self.features = features.split(",")
        self._ids = sender_obs.feature_ids(self.features)
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self._dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.strict_rng = bool(strict_rng)
        self._variant = bool(USE_CWND or USE_LATENCY_NOISE)
        self._use_cwnd = bool(USE_CWND)
        hf = history_len * len(self._ids)
        L = self.L = _lib.load()
        self.h = None
        if self._variant:
            from .multi_env import PccMultiSenderEnv
            self._menv = PccMultiSenderEnv(1, n_senders=1, history_len=history_len, features=features, device=self.device,
                                           ring_capacity=1 << 16, use_cwnd=USE_CWND, use_latency_noise=USE_LATENCY_NOISE)
        else:
            cfg = _lib.PccConfig()
            L.pcc_default_config(C.byref(cfg))
            cfg.device = self._dev_index
            cfg.n_envs = 1
            cfg.history_len = history_len
            cfg.n_features = len(self._ids)
            for i, fid in enumerate(self._ids):
                cfg.feature_ids[i] = fid
            cfg.rng_kind = _lib.PCC_RNG_MT19937
            cfg.ring_capacity = L.pcc_ring_capacity_for(float(MAX_RATE), float(self.min_bw), float(self.max_lat),
                                                        float(1 + int(np.exp(self.max_queue))))
            self.cfg = cfg
            sb, rb = C.c_uint64(), C.c_uint64()
            _lib.check(L.pcc_workspace_bytes(C.byref(cfg), C.byref(sb), C.byref(rb)))
            self._state_ws = torch.empty(sb.value, dtype=torch.uint8, device=self.device)
            self._ring_ws = torch.empty(rb.value, dtype=torch.uint8, device=self.device)
            self.h = C.c_void_p()
            _lib.check(L.pcc_create(C.byref(self.h), C.byref(cfg), self._state_ws.data_ptr(), self._ring_ws.data_ptr()))
            # page-locked host block of the one-call step (pcc_step_host_submit / _wait), laid out like the library's
            # device staging [obs | reward | info | counts | done] so that the results arrive with a single copy
            nbytes = 8 * hf + 8 + 8 * _lib.PCC_INFO_WIDTH + 12 + 1
            self._h_block = torch.zeros(nbytes + 7, dtype=torch.uint8, pin_memory=True)
            self._h_action = torch.zeros(1, dtype=torch.float64, pin_memory=True)
            blk = self._h_block.numpy()
            o_r, o_i = 8 * hf, 8 * hf + 8
            o_c = o_i + 8 * _lib.PCC_INFO_WIDTH
            self._np = {"action": self._h_action.numpy(), "obs": blk[0:o_r].view(np.float64), "reward": blk[o_r:o_i].view(np.float64),
                        "info": blk[o_i:o_c].view(np.float64), "counts": blk[o_c:o_c + 12].view(np.int32),
                        "done": blk[o_c + 12:o_c + 13]}
            self._ptr = {k: v.ctypes.data_as(C.c_void_p) for k, v in self._np.items()}
            self._d_obs = torch.zeros(hf, dtype=torch.float64, device=self.device)
            self._mt = (C.c_uint32 * 625)()
            self._rng_on_device = False     # True: the device holds the newest MT19937 state (lazy mode)
            self._gauss_next = None

        self.links = None
        self.senders = None
        self.create_new_links_and_senders()   # the reference draws one throw-away link here (:366)
        self.run_dur = None
        self.run_period = 0.1
        self.steps_taken = 0
        self.max_steps = MAX_STEPS
        if USE_CWND:                            # :376-379
            self.action_space = spaces.Box(np.array([-1e12, -1e12]), np.array([1e12, 1e12]), dtype=np.float32)
        else:
            self.action_space = spaces.Box(np.array([-1e12]), np.array([1e12]), dtype=np.float32)
        single_obs_min_vec = sender_obs.get_min_obs_vector(self.features)
        single_obs_max_vec = sender_obs.get_max_obs_vector(self.features)
        self.observation_space = spaces.Box(np.tile(single_obs_min_vec, self.history_len),
                                            np.tile(single_obs_max_vec, self.history_len),
                                            dtype=np.float32)
        self.reward_sum = 0.0
        self.reward_ewma = 0.0
        self.event_record = {"Events": []}
        self.episodes_run = -1
        self.last_counts = (0, 0, 0)

    # -- global `random` <-> device MT19937 ------------------------------------------------
    def _push_rng(self):
        st = random.getstate()
        self._mt[:] = st[1]
        _lib.check(self.L.pcc_set_mt_state(self.h, 0, self._mt))
        self._gauss_next = st[2]

    def _pull_rng(self):
        _lib.check(self.L.pcc_get_mt_state(self.h, 0, self._mt))
        random.setstate((3, tuple(self._mt), self._gauss_next))
        self._rng_on_device = False

    def sync_rng(self):
        """Lazy mode: brings Python's global `random` up to date with the draws the device has made since the last reset
        (call it before drawing from `random` yourself if you need the reference's exact stream)."""
        if self.h and self._rng_on_device:
            self._pull_rng()

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def seed(self, seed=None):
        self.rand, seed = seeding.np_random(seed)   # as in the reference: seeds an unused RNG (:396-398)
        return [seed]

    def create_new_links_and_senders(self):
        """The five draws of gym/network_sim.py:455-466 from the global `random`, host side."""
        bw = random.uniform(self.min_bw, self.max_bw)
        lat = random.uniform(self.min_lat, self.max_lat)
        queue = 1 + int(np.exp(random.uniform(self.min_queue, self.max_queue)))
        loss = random.uniform(self.min_loss, self.max_loss)
        start_rate = random.uniform(0.3, 1.5) * bw
        self.link_params = dict(bw=bw, lat=lat, queue=queue, loss=loss, start_rate=start_rate)
        self.run_dur = 3 * lat

    def reset(self):
        torch = self.torch
        self.steps_taken = 0
        if not self._variant:
            self.sync_rng()                     # the five draws below continue the stream the device left
        self.create_new_links_and_senders()
        self.episodes_run += 1
        if self.episodes_run > 0 and self.episodes_run % 100 == 0:
            self.dump_events_to_file("pcc_env_log_run_%d.json" % self.episodes_run)
        self.event_record = {"Events": []}
        p = self.link_params
        self.reward_ewma *= 0.99
        self.reward_ewma += 0.01 * self.reward_sum
        self.reward_sum = 0.0
        if self._variant:
            self._menv.seed(seeds=np.array([random.getrandbits(64)], dtype=np.uint64))
            obs = self._menv.reset(dict(bw=[p["bw"]], lat=[p["lat"]], queue=[p["queue"]], loss=[p["loss"]]), [[p["start_rate"]]])
            self._menv.check()
            return obs.reshape(-1).cpu().numpy()
        f = lambda v, dt=torch.float64: torch.tensor([v], dtype=dt, device=self.device)
        bw, lat, q = f(p["bw"]), f(p["lat"]), f(p["queue"], torch.int64)
        loss, rate = f(p["loss"]), f(p["start_rate"])
        self._push_rng()
        _lib.check(self.L.pcc_reset(self.h, None, bw.data_ptr(), lat.data_ptr(), q.data_ptr(), loss.data_ptr(),
                                    rate.data_ptr(), self._d_obs.data_ptr(), self._stream()))
        if self.strict_rng:
            self._pull_rng()   # synchronises
        else:
            self._rng_on_device = True
        _lib.check(self.L.pcc_check(self.h, self._stream()))   # synchronises; a ring overflow raises here
        return self._d_obs.cpu().numpy()

    def step(self, actions):
        a = np.asarray(actions, dtype=np.float64).reshape(-1)       # float64 first (hard part 9)
        if self._variant:
            return self._step_variant(a)
        self._np["action"][0] = a[0]
        if self.strict_rng:
            self._push_rng()
        t = C.c_int64()
        P = self._ptr
        _lib.check(self.L.pcc_step_host_submit(self.h, P["action"], P["obs"], P["reward"], P["done"], P["counts"], P["info"],
                                               self._stream(), C.byref(t)))
        _lib.check(self.L.pcc_step_host_wait(self.h, t.value))     # the one synchronisation of a step
        if self.strict_rng:
            self._pull_rng()
        else:
            self._rng_on_device = True
        self.steps_taken += 1
        info = self._np["info"]
        reward = np.float64(self._np["reward"][0])
        c = self._np["counts"]
        self.last_counts = (int(c[0]), int(c[1]), int(c[2]))
        event = {"Name": "Step", "Time": self.steps_taken, "Reward": float(reward),
                 "Send Rate": info[0], "Throughput": info[1], "Latency": info[2], "Loss Rate": info[3],
                 "Latency Inflation": info[4], "Latency Ratio": info[5], "Send Ratio": info[6]}
        self.event_record["Events"].append(event)
        self.run_dur = info[10]
        self.cur_time = info[8]
        self.rate = info[9]
        self.reward_sum += reward
        done = self.steps_taken >= self.max_steps
        if done or self.steps_taken % 64 == 0:
            _lib.check(self.L.pcc_check(self.h, self._stream()))   # a ring overflow must not stay silent
        return self._np["obs"].copy(), reward, done, {}

    def _step_variant(self, a):
        """USE_CWND / USE_LATENCY_NOISE: the per-env event-heap engine, one sender (network_sim.py:409-414)."""
        if self._use_cwnd:
            obs, r, d, info = self._menv.step([[a[0]]], [[a[1]]])
        else:
            obs, r, d, info = self._menv.step([[a[0]]])
        self.steps_taken += 1
        reward = np.float64(r.reshape(-1)[0].item())
        self.last_counts = tuple(int(x) for x in info["counts"].reshape(-1).cpu().numpy())
        self.event_record["Events"].append({"Name": "Step", "Time": self.steps_taken, "Reward": float(reward)})
        self.reward_sum += reward
        done = self.steps_taken >= self.max_steps
        if done:
            self._menv.check()
        return obs.reshape(-1).cpu().numpy(), reward, done, {}

    def render(self, mode='human'):
        pass

    def close(self):
        if getattr(self, "h", None):
            try:
                self.sync_rng()
            except Exception:
                pass
            self.L.pcc_destroy(self.h)
            self.h = None
        if getattr(self, "_menv", None) is not None:
            self._menv.close()
            self._menv = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def dump_events_to_file(self, filename):
        with open(filename, 'w') as f:
            json.dump(self.event_record, f, indent=4)


if gym is not None:
    try:
        register(id='PccNs-v0', entry_point=__name__ + ':SimulatedNetworkEnv')   # whichever name this module was imported under
    except Exception:
        pass
