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

try:  # package import or top-level `import network_sim` (sys.path drop-in)
    from . import _lib, sender_obs
except ImportError:  # pragma: no cover
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

# reference constants (gym/network_sim.py:33-54)
MAX_RATE = 1000
MIN_RATE = 40
REWARD_SCALE = 0.001
MAX_STEPS = 400
BYTES_PER_PACKET = 1500
# The reference's module switches (network_sim.py:51-54).  This drop-in class runs the shipped configuration
# (both False) on the three-cursor kernels with Python's own MT19937 stream.  The two variants exist on the GPU
# too -- PccMultiSenderEnv(n, n_senders=1, use_cwnd=True, use_latency_noise=True), Philox streams -- but not
# behind this class: setting a switch here raises instead of silently simulating something else.
USE_CWND = False
USE_LATENCY_NOISE = False


class SimulatedNetworkEnv(_EnvBase):

    def __init__(self, history_len=10,
                 features="sent latency inflation,latency ratio,send ratio", device=None):
        import torch
        if USE_CWND or USE_LATENCY_NOISE:
            raise NotImplementedError("USE_CWND / USE_LATENCY_NOISE: use pcc_rl_b200.PccMultiSenderEnv(n, n_senders=1, "
                                      "use_cwnd=..., use_latency_noise=...) -- the drop-in env runs the shipped configuration")
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

        L = self.L = _lib.load()
        cfg = _lib.PccConfig()
        L.pcc_default_config(C.byref(cfg))
        cfg.device = self.device.index
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
        hf = history_len * len(self._ids)
        pin = dict(pin_memory=True)
        self._h_action = torch.zeros(1, dtype=torch.float64, **pin)
        self._h_obs = torch.zeros(hf, dtype=torch.float64, **pin)
        self._h_reward = torch.zeros(1, dtype=torch.float64, **pin)
        self._h_done = torch.zeros(1, dtype=torch.uint8, **pin)
        self._h_counts = torch.zeros(3, dtype=torch.int32, **pin)
        self._d_info = torch.zeros(_lib.PCC_INFO_WIDTH, dtype=torch.float64, device=self.device)
        self._d_obs = torch.zeros(hf, dtype=torch.float64, device=self.device)
        self._d_scal = torch.zeros(1, dtype=torch.float64, device=self.device)
        self._d_done = torch.zeros(1, dtype=torch.uint8, device=self.device)
        self._d_counts = torch.zeros(3, dtype=torch.int32, device=self.device)
        self._d_action = torch.zeros(1, dtype=torch.float64, device=self.device)
        self._mt = (C.c_uint32 * 625)()

        self.links = None
        self.senders = None
        self.create_new_links_and_senders()   # the reference draws one throw-away link here (:366)
        self.run_dur = None
        self.run_period = 0.1
        self.steps_taken = 0
        self.max_steps = MAX_STEPS
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
        self.create_new_links_and_senders()
        self.episodes_run += 1
        if self.episodes_run > 0 and self.episodes_run % 100 == 0:
            self.dump_events_to_file("pcc_env_log_run_%d.json" % self.episodes_run)
        self.event_record = {"Events": []}
        p = self.link_params
        f = lambda v, dt=torch.float64: torch.tensor([v], dtype=dt, device=self.device)
        bw, lat, q = f(p["bw"]), f(p["lat"]), f(p["queue"], torch.int64)
        loss, rate = f(p["loss"]), f(p["start_rate"])
        self._push_rng()
        _lib.check(self.L.pcc_reset(self.h, None, bw.data_ptr(), lat.data_ptr(), q.data_ptr(), loss.data_ptr(),
                                    rate.data_ptr(), self._d_obs.data_ptr(), self._stream()))
        self._pull_rng()   # synchronises
        _lib.check(self.L.pcc_check(self.h, self._stream()))
        self.reward_ewma *= 0.99
        self.reward_ewma += 0.01 * self.reward_sum
        self.reward_sum = 0.0
        return self._d_obs.cpu().numpy()

    def step(self, actions):
        action = float(np.asarray(actions, dtype=np.float64).reshape(-1)[0])  # float64 first (hard part 9)
        self._d_action[0] = action
        self._push_rng()
        _lib.check(self.L.pcc_step(self.h, self._d_action.data_ptr(), self._d_obs.data_ptr(),
                                   self._d_scal.data_ptr(), self._d_done.data_ptr(), self._d_counts.data_ptr(),
                                   self._d_info.data_ptr(), self._stream()))
        self._pull_rng()   # synchronises
        self.steps_taken += 1
        sender_obs_arr = self._d_obs.cpu().numpy()
        reward = np.float64(self._d_scal.item())
        info = self._d_info.cpu().numpy()
        self.last_counts = tuple(int(c) for c in self._d_counts.cpu().numpy())
        event = {"Name": "Step", "Time": self.steps_taken, "Reward": float(reward),
                 "Send Rate": info[0], "Throughput": info[1], "Latency": info[2], "Loss Rate": info[3],
                 "Latency Inflation": info[4], "Latency Ratio": info[5], "Send Ratio": info[6]}
        self.event_record["Events"].append(event)
        self.run_dur = info[10]
        self.cur_time = info[8]
        self.rate = info[9]
        self.reward_sum += reward
        return sender_obs_arr, reward, (self.steps_taken >= self.max_steps), {}

    def render(self, mode='human'):
        pass

    def close(self):
        if getattr(self, "h", None):
            self.L.pcc_destroy(self.h)
            self.h = None

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
        register(id='PccNs-v0', entry_point='network_sim:SimulatedNetworkEnv')
    except Exception:
        pass
