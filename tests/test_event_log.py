"""the reference's event-log format from the batched env's `info` rows."""
import json
import os

import numpy as np
import pytest

import pcc_rl_b200
import refharness as rh
from golden_util import load_golden


def test_recorder_writes_the_reference_keys_in_order(tmp_path):
    g = load_golden("mt_seed1234")
    rec = pcc_rl_b200.EventRecorder([3])
    n = 8
    for k in range(5):
        reward = np.zeros(n); info = np.zeros((n, 12)); done = np.zeros(n, dtype=np.uint8)
        reward[3] = g["reward"][k]; info[3, :7] = g["info"][k]
        rec.record(reward, info, done)
    path = tmp_path / "pcc_env_log_run_100.json"
    rec.dump(3, str(path))
    data = json.load(open(path))
    assert list(data.keys()) == ["Events"] and len(data["Events"]) == 5
    for k, ev in enumerate(data["Events"]):
        assert tuple(ev.keys()) == pcc_rl_b200.event_log.EVENT_KEYS
        assert ev["Name"] == "Step" and ev["Time"] == k + 1 and ev["Reward"] == g["reward"][k]
        assert [ev[c] for c in ("Send Rate", "Throughput", "Latency", "Loss Rate", "Latency Inflation", "Latency Ratio",
                                "Send Ratio")] == list(g["info"][k])
    # what gym/graph_run.py:27-34 does with the file
    time_data = [float(e["Time"]) for e in data["Events"][1:]]
    thpt = [float(e["Throughput"]) for e in data["Events"][1:]]
    assert time_data == [2.0, 3.0, 4.0, 5.0] and len(thpt) == 4


def test_recorder_restarts_at_episode_end(tmp_path):
    rec = pcc_rl_b200.EventRecorder([0, 1])
    for k in range(3):
        done = np.array([k == 1, False])
        rec.record(np.array([1.0 * k, 2.0 * k]), np.ones((2, 7)) * k, done)
    assert [e["Time"] for e in rec.records[0]["Events"]] == [1] and [e["Time"] for e in rec.records[1]["Events"]] == [1, 2, 3]
    rec.dump(0, str(tmp_path / "a.json"), finished=True)
    assert len(json.load(open(tmp_path / "a.json"))["Events"]) == 2 and rec.episodes[0] == 1


@pytest.mark.reference
@pytest.mark.skipif(not rh.reference_available(), reason="reference tree not present")
def test_same_file_as_the_reference_writes(tmp_path):
    """random.seed(1234); the reference env; 6 steps; dump_events_to_file -- against the recorder fed with the
    committed outputs of that very run (tests/golden/mt_seed1234.npz)."""
    import random
    ns = rh.load_reference()
    g = load_golden("mt_seed1234")
    with rh.quiet_tmp_cwd():
        random.seed(1234)
        env = ns.SimulatedNetworkEnv()
        env.reset()
        for k in range(6):
            env.step([float(g["action"][k])])
        ref_path = os.path.join(os.getcwd(), "ref.json")
        env.dump_events_to_file(ref_path)
        ref_text = open(ref_path).read()
    rec = pcc_rl_b200.EventRecorder([0])
    for k in range(6):
        rec.record(np.array([g["reward"][k]]), g["info"][k][None, :])
    ours = tmp_path / "ours.json"
    rec.dump(0, str(ours))
    assert open(ours).read() == ref_text
