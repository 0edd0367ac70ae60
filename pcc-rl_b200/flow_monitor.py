"""MI-sample ingestion on the GPU ("flow monitor"): the reference's other producer of monitor intervals.

In the reference, the PCC sender (C++) measures an MI on a REAL flow and calls into Python:
`give_sample(flow_id, bytes_sent, ..., rtt_samples, packet_size, utility)` builds a SenderMonitorInterval,
pushes it into the flow's SenderHistory, and `get_rate(flow_id)` feeds the history array to the agent and
applies its action to the sending rate (udt-plugins/testing/loaded_client.py:43-183); during online training
the same record travels as a text line to ShimNetworkEnv.step (udt-plugins/training/shim.py:31-42,
gym/online/shim_env.py:102-139).  One Python object per flow, one record at a time.

Here:
  * PccFlowMonitor      many flows on one GPU; a batch of records (structure of arrays + CSR sample lists,
                        torch.cuda tensors) -> metrics -> histories -> observations in ONE kernel launch
                        (libpcc_b200.so: pcc_flows_*); rates and the optional MLP agent stay on the device.
  * init / give_sample / get_rate / reset (module level) + PccGymDriver
                        the reference module's own API (same names, argument order and meaning), each call
                        a batch of one -- so the C++ side that embeds loaded_client.py can load this module
                        instead.  The agent is any callable obs -> action (the reference loads a TensorFlow
                        saved model; TF is not a dependency here) or an on-device MLP (set_policy).
  * parse_sample_line / format_sample_line   the shim's wire format.
Metric values are bit-identical to common/sender_obs.py (tests/test_gpu_flows.py replays the reference's
own outputs).  There is no CPU fallback.
"""
import ctypes as C
import random

import numpy as np

from . import _lib, sender_obs

# rate-control constants of the two callers
CLIENT_DEFAULTS = dict(delta_scale=0.05, min_rate=0.5, max_rate=300.0, rate_style=_lib.PCC_RATE_CLIENT)   # loaded_client.py:33-35
SHIM_DEFAULTS = dict(delta_scale=0.025, min_rate=0.25, max_rate=1000.0, rate_style=_lib.PCC_RATE_SHIM)    # shim_env.py:38-44
RESET_RATE_MIN = 6.0     # loaded_client.py:40-41
RESET_RATE_MAX = 6.0
STARTING_RATE = 2.0      # shim_env.py:41

_FIELDS_I64 = ("bytes_sent", "bytes_acked", "bytes_lost", "packet_size")
_FIELDS_F64 = ("send_start", "send_end", "recv_start", "recv_end")


class PccFlowMonitor(object):
    def __init__(self, n_flows, history_len=10, features=sender_obs.DEFAULT_FEATURES, device=None,
                 delta_scale=0.05, min_rate=0.5, max_rate=300.0, rate_style=_lib.PCC_RATE_CLIENT, start_rate=0.0,
                 workspace=None):
        """workspace: a uint8 CUDA tensor holding the state of another monitor of the same shape (its `.workspace`,
        e.g. restored with torch.load): the new monitor adopts it as is (pcc_flows_attach) -- checkpoint / resume."""
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("pcc_rl_b200.PccFlowMonitor needs a CUDA device; there is no CPU fallback")
        self.torch, self.L = torch, _lib.load()
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.n_flows = int(n_flows)
        self.history_len = int(history_len)
        self.features = sender_obs.feature_names(features)
        self.feature_ids = sender_obs.feature_ids(features)
        self.obs_dim = self.history_len * len(self.feature_ids)
        cfg = _lib.PccFlowsConfig()
        self.L.pcc_flows_default_config(C.byref(cfg))
        cfg.device = self.device.index if self.device.index is not None else torch.cuda.current_device()
        cfg.n_flows, cfg.history_len, cfg.n_features = self.n_flows, self.history_len, len(self.feature_ids)
        for i, fid in enumerate(self.feature_ids):
            cfg.feature_ids[i] = fid
        cfg.delta_scale, cfg.min_rate, cfg.max_rate, cfg.rate_style = delta_scale, min_rate, max_rate, rate_style
        self.cfg = cfg
        nb = C.c_uint64()
        _lib.check(self.L.pcc_flows_workspace_bytes(C.byref(cfg), C.byref(nb)))
        with torch.cuda.device(self.device):
            self.h = C.c_void_p()
            if workspace is not None:
                if workspace.dtype != torch.uint8 or workspace.numel() != nb.value or not workspace.is_cuda:
                    raise ValueError("workspace must be a uint8 CUDA tensor of %d bytes" % nb.value)
                self.workspace = workspace.to(self.device).contiguous()
                _lib.check(self.L.pcc_flows_attach(C.byref(self.h), C.byref(cfg), self.workspace.data_ptr()))
            else:
                self.workspace = torch.empty(nb.value, dtype=torch.uint8, device=self.device)   # = the checkpoint
                _lib.check(self.L.pcc_flows_create(C.byref(self.h), C.byref(cfg), self.workspace.data_ptr()))
        self._keep = None
        self._policy = None
        if start_rate and workspace is None:
            self.set_rates(rate=start_rate)

    # -- plumbing ---------------------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def close(self):
        if getattr(self, "h", None):
            self.L.pcc_flows_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self):
        return int(self.L.pcc_flows_launch_count(self.h))

    def _dev(self, a, dtype):
        torch = self.torch
        if isinstance(a, torch.Tensor):
            return a.to(self.device, dtype).contiguous()
        return torch.as_tensor(np.ascontiguousarray(a)).to(self.device, dtype)

    def _mask(self, mask):
        if mask is None:
            return None
        return self._dev(mask, self.torch.uint8)

    def make_batch(self, flow, bytes_sent, bytes_acked, bytes_lost, send_start, send_end, recv_start, recv_end,
                   packet_size, rtt_off, rtt):
        """Moves a batch (numpy arrays or tensors) to the device; returns the dict give_samples takes."""
        t = self.torch
        b = dict(flow=self._dev(flow, t.int32), bytes_sent=self._dev(bytes_sent, t.int64),
                 bytes_acked=self._dev(bytes_acked, t.int64), bytes_lost=self._dev(bytes_lost, t.int64),
                 packet_size=self._dev(packet_size, t.int64), send_start=self._dev(send_start, t.float64),
                 send_end=self._dev(send_end, t.float64), recv_start=self._dev(recv_start, t.float64),
                 recv_end=self._dev(recv_end, t.float64), rtt_off=self._dev(rtt_off, t.int64),
                 rtt=self._dev(rtt, t.float64))
        assert b["rtt_off"].numel() == b["flow"].numel() + 1
        return b

    # -- the path ---------------------------------------------------------------------------------------
    def give_samples(self, batch, unique_flows=False, want_obs=True, want_metrics=False, obs_out=None):
        """Ingests a batch of MI records (give_sample x R).  Returns (obs[R, H*F] or None, metrics[R, 12] or
        None): each record's flow observation right after it, and its 12 raw metric values."""
        torch = self.torch
        R = int(batch["flow"].numel())
        mb = _lib.PccMiBatch()
        mb.n_records = R
        mb.flow = batch["flow"].data_ptr()
        for k in _FIELDS_I64 + _FIELDS_F64:
            setattr(mb, k, batch[k].data_ptr())
        mb.rtt_offsets = batch["rtt_off"].data_ptr()
        mb.rtt_samples = batch["rtt"].data_ptr()
        obs = metrics = None
        with torch.cuda.device(self.device):
            if want_obs:
                obs = obs_out if obs_out is not None else torch.empty((R, self.obs_dim), dtype=torch.float64,
                                                                      device=self.device)
            if want_metrics:
                metrics = torch.empty((R, _lib.PCC_N_METRICS), dtype=torch.float64, device=self.device)
            _lib.check(self.L.pcc_flows_give_samples(self.h, C.byref(mb), 1 if unique_flows else 0,
                                                     obs.data_ptr() if obs is not None else None,
                                                     metrics.data_ptr() if metrics is not None else None,
                                                     self._stream()))
        self._keep = batch
        return obs, metrics

    def obs(self):
        """history.as_array() of every flow: [n_flows, H*F], oldest MI first."""
        o = self.torch.empty((self.n_flows, self.obs_dim), dtype=self.torch.float64, device=self.device)
        _lib.check(self.L.pcc_flows_get_obs(self.h, o.data_ptr(), self._stream()))
        return o

    def reset(self, mask=None, mode=_lib.PCC_FLOW_RESET_CLIENT):
        m = self._mask(mask)
        _lib.check(self.L.pcc_flows_reset(self.h, m.data_ptr() if m is not None else None, int(mode), self._stream()))
        self._keep_m = m

    def set_rates(self, rates=None, rate=0.0, mask=None):
        m = self._mask(mask)
        r = self._dev(rates, self.torch.float64) if rates is not None else None
        _lib.check(self.L.pcc_flows_set_rates(self.h, m.data_ptr() if m is not None else None,
                                              r.data_ptr() if r is not None else None, float(rate), self._stream()))
        self._keep_r = (m, r)

    def get_rates(self, actions=None, mask=None):
        """Applies actions[flow] (optional) to the selected flows and returns every flow's rate [n_flows]."""
        a = self._dev(actions, self.torch.float64) if actions is not None else None
        m = self._mask(mask)
        out = self.torch.empty(self.n_flows, dtype=self.torch.float64, device=self.device)
        _lib.check(self.L.pcc_flows_get_rates(self.h, a.data_ptr() if a is not None else None,
                                              m.data_ptr() if m is not None else None, out.data_ptr(), self._stream()))
        self._keep_a = (a, m)
        return out

    def set_policy(self, w1, b1, w2, b2, w3, b3):
        """The agent as an on-device MLP (tanh hidden layers, linear output; stable_solve.py:30-45)."""
        t = self.torch
        ws = [self._dev(x, t.float64) for x in (w1, b1, w2, b2, w3, b3)]
        pol = _lib.PccPolicy()
        pol.w1, pol.b1, pol.w2, pol.b2, pol.w3, pol.b3 = (x.data_ptr() for x in ws)
        pol.n_in, pol.h1, pol.h2 = self.obs_dim, ws[1].numel(), ws[3].numel()
        assert ws[0].numel() == pol.n_in * pol.h1 and ws[2].numel() == pol.h1 * pol.h2 and ws[4].numel() == pol.h2
        self._policy = (pol, ws)

    def act(self):
        """actions[flow] = policy(obs[flow]) on the device (deterministic)."""
        if self._policy is None:
            raise RuntimeError("set_policy first")
        out = self.torch.empty(self.n_flows, dtype=self.torch.float64, device=self.device)
        _lib.check(self.L.pcc_flows_act(self.h, C.byref(self._policy[0]), out.data_ptr(), self._stream()))
        return out

    def column(self, name):
        out = self.torch.empty(self.n_flows, dtype=self.torch.float64, device=self.device)
        _lib.check(self.L.pcc_flows_get_column(self.h, name.encode(), out.data_ptr(), self._stream()))
        return out

    def check(self):
        _lib.check(self.L.pcc_flows_check(self.h, self._stream()))


# ---------------------------------------------------------------------------------------------------------
# The shim's wire format (udt-plugins/training/shim.py:31-42 writes it, gym/online/shim_env.py:108-121 reads it)
# ---------------------------------------------------------------------------------------------------------
def format_sample_line(flow_id, bytes_sent, bytes_acked, bytes_lost, send_start_time, send_end_time,
                       recv_start_time, recv_end_time, rtt_samples, packet_size, utility):
    return "%d;%d;%d;%d;%f;%f;%f;%f;%s;%d;%f\n" % (flow_id, bytes_sent, bytes_acked, bytes_lost, send_start_time,
                                                    send_end_time, recv_start_time, recv_end_time,
                                                    list(rtt_samples), packet_size, utility)


def parse_sample_line(data):
    """Takes what conn.recv() returned (possibly several lines): like the reference, the LAST complete line."""
    import ast
    vals = data.split("\n")[-2].split(";")
    return dict(flow_id=int(vals[0]), bytes_sent=int(vals[1]), bytes_acked=int(vals[2]), bytes_lost=int(vals[3]),
                send_start_time=float(vals[4]), send_end_time=float(vals[5]), recv_start_time=float(vals[6]),
                recv_end_time=float(vals[7]), rtt_samples=[float(r) for r in ast.literal_eval(vals[8])],
                packet_size=int(vals[9]), utility=float(vals[10]))


# ---------------------------------------------------------------------------------------------------------
# loaded_client.py's module API (udt-plugins/testing/loaded_client.py:43-183), on top of one PccFlowMonitor
# ---------------------------------------------------------------------------------------------------------
MAX_FLOWS = 1024
_monitor = None
_agent_factory = None
_history_len = 10
_features = sender_obs.DEFAULT_FEATURES


def configure(agent_factory=None, history_len=10, features=sender_obs.DEFAULT_FEATURES, max_flows=MAX_FLOWS):
    """agent_factory() -> object with act(obs) and reset() (the reference constructs
    loaded_agent.LoadedModelAgent(MODEL_PATH) per flow, loaded_client.py:62)."""
    global _monitor, _agent_factory, _history_len, _features, MAX_FLOWS
    _agent_factory, _history_len, _features, MAX_FLOWS = agent_factory, history_len, features, max_flows
    _monitor = None
    PccGymDriver.flow_lookup = {}
    PccGymDriver.next_slot = 0


def _get_monitor():
    global _monitor
    if _monitor is None:
        _monitor = PccFlowMonitor(MAX_FLOWS, _history_len, _features, **CLIENT_DEFAULTS)
    return _monitor


class PccGymDriver(object):
    flow_lookup = {}
    next_slot = 0          # slots are handed out once and never shared: a flow id keeps its slot across re-inits

    def __init__(self, flow_id):
        self.id = flow_id
        self.mon = _get_monitor()
        prev = PccGymDriver.flow_lookup.get(flow_id)
        if prev is not None:
            # init(flow_id) for an id that exists (loaded_client.py:173-175 replaces the dict entry): same GPU slot, fresh
            # history -- and, like the reference's module-level _conn_min_latencies, the flow's connection-min entry stays
            self.slot, mode = prev.slot, _lib.PCC_FLOW_RESET_CLIENT
        else:
            self.slot, mode = PccGymDriver.next_slot, _lib.PCC_FLOW_RESET_NEW
            if self.slot >= self.mon.n_flows:
                raise RuntimeError("more than %d flows: raise max_flows in configure()" % self.mon.n_flows)
            PccGymDriver.next_slot += 1
        self.rate = random.uniform(RESET_RATE_MIN, RESET_RATE_MAX)      # :51
        self.history_len = self.mon.history_len
        self.features = self.mon.features
        self.got_data = False
        self.agent = _agent_factory() if _agent_factory is not None else None
        self._sel = np.zeros(self.mon.n_flows, dtype=np.uint8)
        self._sel[self.slot] = 1
        self.mon.reset(mask=self._sel, mode=mode)
        self.mon.set_rates(rate=self.rate, mask=self._sel)
        PccGymDriver.flow_lookup[flow_id] = self

    def _obs(self):
        return self.mon.obs()[self.slot].cpu().numpy()

    def get_rate(self):                                                 # :72-76
        if self.has_data():
            rate_delta = float(self.agent.act(self._obs()))
            acts = np.zeros(self.mon.n_flows)
            acts[self.slot] = rate_delta
            self.rate = float(self.mon.get_rates(actions=acts, mask=self._sel)[self.slot].item())
        return self.rate * 1e6

    def has_data(self):
        return self.got_data

    def set_current_rate(self, new_rate):
        self.current_rate = new_rate

    def reset_rate(self):                                               # :94-95 (sets an attribute nothing reads)
        self.current_rate = random.uniform(RESET_RATE_MIN, RESET_RATE_MAX)

    def reset_history(self):                                            # :97-101
        self.mon.reset(mask=self._sel, mode=_lib.PCC_FLOW_RESET_CLIENT)
        self.got_data = False

    def reset(self):                                                    # :103-106
        if self.agent is not None:
            self.agent.reset()
        self.reset_rate()
        self.reset_history()

    def give_sample(self, bytes_sent, bytes_acked, bytes_lost, send_start_time, send_end_time, recv_start_time,
                    recv_end_time, rtt_samples, packet_size, utility):  # :111-127
        rtt = np.asarray(rtt_samples, dtype=np.float64)
        b = self.mon.make_batch([self.slot], [bytes_sent], [bytes_acked], [bytes_lost], [send_start_time],
                                [send_end_time], [recv_start_time], [recv_end_time], [packet_size],
                                [0, rtt.size], rtt)
        self.mon.give_samples(b, unique_flows=True, want_obs=False)
        self.got_data = True

    @staticmethod
    def get_by_flow_id(flow_id):
        return PccGymDriver.flow_lookup[flow_id]


def give_sample(flow_id, bytes_sent, bytes_acked, bytes_lost, send_start_time, send_end_time, recv_start_time,
                recv_end_time, rtt_samples, packet_size, utility):     # :129-136
    PccGymDriver.get_by_flow_id(flow_id).give_sample(bytes_sent, bytes_acked, bytes_lost, send_start_time,
                                                     send_end_time, recv_start_time, recv_end_time, rtt_samples,
                                                     packet_size, utility)


def reset(flow_id):                                                     # :170-172
    PccGymDriver.get_by_flow_id(flow_id).reset()


def get_rate(flow_id):                                                  # :174-177
    return PccGymDriver.get_by_flow_id(flow_id).get_rate()


def init(flow_id):                                                      # :179-180
    PccGymDriver(flow_id)
