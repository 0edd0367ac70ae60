"""Drop-in for the reference's gym/online/shim_env.py: `ShimNetworkEnv`, the gym env whose step sends a sending rate
to a real PCC/UDT sender over TCP (localhost:9787) and gets one monitor-interval record back as a `;`-separated text
line (udt-plugins/training/shim.py:31-42 writes it, shim_env.py:102-139 reads it).  Same constructor defaults, same
constants, same wire protocol, same step / reset results -- the feature code (common/sender_obs.py) runs on the GPU
through a one-flow `PccFlowMonitor` (csrc/pcc_flows.cuh), bit-identical to the reference's SenderHistory
(tests/test_gpu_flows.py), and so does the rate update (the monitor's shim rate style: shim_env.py:82-96).

Two parts, so that the socket side can be tested without a GPU:
  * ShimLink          the TCP side: bind / accept lazily, send the rate, read one record.
  * ShimNetworkEnv    the env: ShimLink + PccFlowMonitor(1 flow).
Difference from the reference, on purpose: a record is read until its newline instead of with a single recv(1024) --
a reference record with more than ~90 RTT samples does not fit 1024 bytes and is cut there.  Records that fit are
handled identically (the last complete line of what arrived is used, shim_env.py:108-109).
"""
import socket

import numpy as np

from . import _lib, sender_obs
from .flow_monitor import PccFlowMonitor, SHIM_DEFAULTS, STARTING_RATE, parse_sample_line
from .network_sim import _EnvBase, arg_or_default, gym, spaces, seeding

RESET_INTERVAL = 400        # shim_env.py:36
MAX_RATE = 1000.0           # Mbit/s, :39-41
MIN_RATE = 0.25
DELTA_SCALE = 0.025         # :43
SHIM_PORT = 9787            # :63


class ShimLink(object):
    """The env's end of the shim connection (shim_env.py:61-63, 102-108)."""

    def __init__(self, host="localhost", port=SHIM_PORT):
        self.sock = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        self.sock.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        self.sock.setblocking(1)
        self.sock.bind((host, port))
        # the reference starts listening in its first step(); listening from the start only means that a sender which
        # connects earlier waits in the backlog instead of being refused
        self.sock.listen()
        self.port = self.sock.getsockname()[1]
        self.conn, self.addr = None, None

    def accept(self):
        if self.conn is None:
            print("Listening for connection from network sender")
            self.conn, self.addr = self.sock.accept()

    def exchange(self, rate):
        """Sends the rate (`str(float)`, as the reference does) and returns the parsed record that comes back."""
        self.accept()
        self.conn.send(str(rate).encode())
        data = self.conn.recv(1024).decode()
        while not data.endswith("\n"):
            more = self.conn.recv(65536).decode()
            if not more:
                raise ConnectionError("the network sender closed the shim connection in the middle of a record")
            data += more
        return parse_sample_line(data)

    def close(self):
        for s in (self.conn, self.sock):
            try:
                if s is not None:
                    s.close()
            except OSError:
                pass
        self.conn = None


class ShimNetworkEnv(_EnvBase):
    def __init__(self, history_len=arg_or_default("--history-len", default=10),
                 features=arg_or_default("--input-features", default=sender_obs.DEFAULT_FEATURES),
                 host="localhost", port=SHIM_PORT, device=None):
        self.viewer = None
        self.rand = None
        self.link = ShimLink(host, port)
        self.features = sender_obs.feature_names(features)
        self.history_len = int(history_len)
        # one flow (sender id 0, shim_env.py:69), the shim's rate constants, history on the device
        self.mon = PccFlowMonitor(1, self.history_len, features, device=device, start_rate=STARTING_RATE, **SHIM_DEFAULTS)
        self.rate = STARTING_RATE
        lo = np.tile(sender_obs.get_min_obs_vector(self.features), self.history_len)
        hi = np.tile(sender_obs.get_max_obs_vector(self.features), self.history_len)
        self.observation_space = spaces.Box(lo, hi, dtype=np.float32)
        self.action_space = spaces.Box(np.array([-1e12]), np.array([1e12]), dtype=np.float32)
        self.steps_taken = 0
        self.reward_sum = 0.0
        self.reward_ewma = 0.0

    # -- rate control (shim_env.py:82-96), on the device: delta = action * 0.025; rate * (1 + delta) or rate / (1 - delta);
    # clamped to [0.25, 1000]
    def apply_action(self, action):
        self.rate = float(self.mon.get_rates(actions=[float(action)])[0].item())

    def set_rate(self, new_rate):
        self.rate = min(max(float(new_rate), MIN_RATE), MAX_RATE)
        self.mon.set_rates(rate=self.rate)

    def seed(self, seed=None):
        self.rand, seed = seeding.np_random(seed)
        return [seed]

    def _obs(self):
        return self.mon.obs()[0].cpu().numpy()

    def step(self, action):
        self.apply_action(np.asarray(action, dtype=np.float64).reshape(-1)[0])
        rec = self.link.exchange(self.rate)                              # :106-121
        rtt = np.asarray(rec["rtt_samples"], dtype=np.float64)
        b = self.mon.make_batch([0], [rec["bytes_sent"]], [rec["bytes_acked"]], [rec["bytes_lost"]],
                                [rec["send_start_time"]], [rec["send_end_time"]], [rec["recv_start_time"]],
                                [rec["recv_end_time"]], [rec["packet_size"]], [0, rtt.size], rtt)
        obs, _ = self.mon.give_samples(b, unique_flows=True)             # :123-134
        rew = rec["utility"]
        self.reward_sum += rew
        self.steps_taken += 1
        done = self.steps_taken > RESET_INTERVAL                         # :137
        return obs[0].cpu().numpy(), rew, done, {}

    def reset(self):
        # a fresh SenderHistory for sender 0 (:141): the module-level connection-min entry of the reference survives, the
        # new empty MIs do not see it (PCC_FLOW_RESET_SHIM, include/pcc_b200.h)
        self.mon.reset(mode=_lib.PCC_FLOW_RESET_SHIM)
        self.reward_ewma *= 0.99
        self.reward_ewma += 0.01 * self.reward_sum
        print("Reward: %0.2f, Ewma Reward: %0.2f" % (self.reward_sum, self.reward_ewma))
        self.reward_sum = 0.0
        self.steps_taken = 0
        self.set_rate(STARTING_RATE)
        return self._obs()

    def render(self, mode='human'):
        pass

    def close(self):
        self.link.close()
        self.mon.close()
        if self.viewer:
            self.viewer.close()
            self.viewer = None


if gym is not None:
    from gym.envs.registration import register
    register(id='NetShim-v0', entry_point=__name__ + ':ShimNetworkEnv')
