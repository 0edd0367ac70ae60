"""The reference's event-log format (`pcc_env_log_run_N.json`) for the batched env.

SimulatedNetworkEnv.step appends one dict per monitor interval to `event_record["Events"]` (gym/network_sim.py:421-436)
and `dump_events_to_file` writes `json.dump(event_record, f, indent=4)` every 100 episodes (:475-476, 494-496);
gym/graph_run.py:27-34 plots Time, Reward, Send Rate, Throughput, Latency, Loss Rate from it.  The drop-in
`network_sim.SimulatedNetworkEnv` keeps that record itself.  For `PccBatchEnv` this recorder does the same for a chosen
set of envs from the `info` rows of `step` (columns 0-6 of PCC_INFO_WIDTH are exactly the seven logged metrics), so the
reference's plotting script works on a batched run unchanged.
"""
import json

import numpy as np

# key order of network_sim.py:423-433
EVENT_KEYS = ("Name", "Time", "Reward", "Send Rate", "Throughput", "Latency", "Loss Rate", "Latency Inflation",
              "Latency Ratio", "Send Ratio")
_INFO_COLUMNS = ("Send Rate", "Throughput", "Latency", "Loss Rate", "Latency Inflation", "Latency Ratio", "Send Ratio")


def make_event(steps_taken, reward, info_row):
    """One entry of event_record["Events"]: `steps_taken` after the step (network_sim.py:419, 424)."""
    ev = {"Name": "Step", "Time": int(steps_taken), "Reward": float(reward)}
    for k, name in enumerate(_INFO_COLUMNS):
        ev[name] = float(info_row[k])
    return ev


class EventRecorder(object):
    """Collects the event records of selected envs of a PccBatchEnv(want_info=True).

        rec = EventRecorder(env_ids=[0, 17])
        obs, reward, done, info = env.step(actions); rec.record(reward, info["metrics"], done)
        rec.dump(0, "pcc_env_log_run_100.json")         # the file gym/graph_run.py reads
    An env's record restarts when it finishes an episode (as the reference's does at reset, :477).
    """

    def __init__(self, env_ids):
        self.env_ids = [int(e) for e in env_ids]
        self.records = {e: {"Events": []} for e in self.env_ids}
        self.steps = {e: 0 for e in self.env_ids}
        self.episodes = {e: 0 for e in self.env_ids}
        self.last_finished = {}          # env id -> record of its last completed episode

    def record(self, reward, info, done=None):
        """reward [N], info [N, >= 7] (tensors or arrays) of one step; done [N] optional."""
        to_np = lambda a: a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)
        idx = np.asarray(self.env_ids)
        if hasattr(reward, "detach"):    # move only the selected rows off the device
            import torch
            tidx = torch.as_tensor(idx, device=reward.device)
            r, m = to_np(reward[tidx]), to_np(info[tidx])
            d = to_np(done[tidx]) if done is not None else None
        else:
            r, m = to_np(reward)[idx], to_np(info)[idx]
            d = to_np(done)[idx] if done is not None else None
        for k, e in enumerate(self.env_ids):
            self.steps[e] += 1
            self.records[e]["Events"].append(make_event(self.steps[e], r[k], m[k]))
            if d is not None and bool(d[k]):
                self.episodes[e] += 1
                self.last_finished[e] = self.records[e]
                self.records[e] = {"Events": []}
                self.steps[e] = 0

    def dump(self, env_id, filename, finished=False):
        """json.dump(event_record, f, indent=4) like dump_events_to_file (:494-496); finished=True writes the last
        completed episode of the env instead of the running one."""
        rec = self.last_finished[int(env_id)] if finished else self.records[int(env_id)]
        with open(filename, "w") as f:
            json.dump(rec, f, indent=4)
