"""-m gpu: the lane-per-env ("packed") execution of the env step (pcc_packed.cuh, pcc_step_packed_kernel) against the
CPU oracle and the reference's golden trajectories, bit for bit.  Big batches select it by default; here it is forced
on small ones (PCC_B200_MODE=packed) with thresholds that send the envs through each of the kernel's three roles -- 32
envs per warp, one per lane; four envs per warp, 8 lanes each; a heavy env alone in its warp -- and through all of them."""
import os

import numpy as np
import pytest

import oracle
import test_gpu_parity as base
from golden_util import golden_names

pytestmark = pytest.mark.gpu


MODES = {   # predicted packets above which an env is run solo / by 8 lanes (quad); re-sort period
    "lanes": dict(PCC_B200_SOLO="65534", PCC_B200_QUAD="65534", PCC_B200_PACKED_EVERY="1"),
    "quad": dict(PCC_B200_SOLO="65534", PCC_B200_QUAD="8", PCC_B200_PACKED_EVERY="1"),
    "mixed": dict(PCC_B200_SOLO="256", PCC_B200_QUAD="48", PCC_B200_PACKED_EVERY="1"),
    "every4": dict(PCC_B200_SOLO="700", PCC_B200_QUAD="100", PCC_B200_PACKED_EVERY="4"),
}


@pytest.fixture(params=sorted(MODES))
def packed_mode(request, monkeypatch):
    monkeypatch.setenv("PCC_B200_MODE", "packed")
    for k, v in MODES[request.param].items():
        monkeypatch.setenv(k, v)
    return request.param


@pytest.mark.parametrize("name", golden_names("philox_"))
def test_packed_matches_reference_golden_philox(packed_mode, name):
    base.test_cuda_matches_reference_golden_philox(name)


def test_packed_auto_reset_multi_episode_all_features(packed_mode):
    base.test_cuda_auto_reset_multi_episode_vs_oracle()


def test_packed_config2_full_episode_vs_oracle(packed_mode):
    base.test_cuda_config2_full_episode_vs_oracle()


def test_packed_ragged_batch_vs_oracle(monkeypatch):
    """1 000 envs (the last warp owns 8), adversarial parameter ranges (tiny queues, heavy loss, long delays), 120 steps."""
    import torch
    monkeypatch.setenv("PCC_B200_MODE", "packed")
    monkeypatch.setenv("PCC_B200_SOLO", "600")
    monkeypatch.setenv("PCC_B200_QUAD", "90")
    n, steps = 1000, 120
    g = np.random.default_rng(21)
    bw = g.uniform(50, 900, n)
    p = dict(bw=bw, lat=g.uniform(0.005, 0.6, n), queue=1 + np.exp(g.uniform(0, 7, n)).astype(np.int64),
             loss=np.where(g.random(n) < 0.2, 0.0, g.uniform(0, 0.4, n)), start_rate=g.uniform(0.2, 2.5, n) * bw)
    seeds = np.arange(n, dtype=np.uint64) * np.uint64(7919) + np.uint64(5)
    acts = g.normal(0, 2.0, (steps, n))
    env = base._env(n_envs=n, auto_reset=False, features=",".join(oracle.METRIC_NAMES), history_len=3)
    env.seed(seeds=seeds)
    env.reset(params=p)
    a_dev = torch.as_tensor(acts, device=env.device)
    rew, cnt = [], []
    for t in range(steps):
        obs, r, d, info = env.step(a_dev[t])
        rew.append(r.clone()); cnt.append(info["counts"].clone())
    env.check()
    ref = oracle.batch_run(p["bw"], p["lat"], p["queue"], p["loss"], p["start_rate"], seeds, steps, actions=acts,
                           n_threads=os.cpu_count() or 1, trajectories=True, history_len=3,
                           features=",".join(oracle.METRIC_NAMES))
    assert np.array_equal(torch.stack(cnt).cpu().numpy(), ref["count_traj"])
    assert np.array_equal(torch.stack(rew).cpu().numpy(), ref["reward_traj"])
    assert np.array_equal(obs.cpu().numpy(), ref["obs"])


@pytest.mark.parametrize("mode", ["warp", "packed"])
def test_tiny_queues_vs_oracle(mode, monkeypatch):
    """Queues of 0, 1 and 2 packets (max_queue_delay == 0 / == 1/bw: tail_drop_threshold's corner; the reset kernel
    must terminate and every step must match the oracle), both step kernels."""
    import torch
    monkeypatch.setenv("PCC_B200_MODE", mode)
    monkeypatch.setenv("PCC_B200_SOLO", "600")
    monkeypatch.setenv("PCC_B200_QUAD", "90")
    n, steps = 384, 60
    g = np.random.default_rng(77)
    bw = np.exp(g.uniform(np.log(40), np.log(5000), n))
    p = dict(bw=bw, lat=np.exp(g.uniform(np.log(0.001), np.log(0.5), n)), queue=g.integers(0, 3, n).astype(np.int64),
             loss=g.choice([0.0, 0.05, 0.3, 1.0], n), start_rate=g.uniform(40, 1000, n))
    seeds = np.arange(n, dtype=np.uint64) * np.uint64(104729) + np.uint64(11)
    acts = g.normal(0, 3.0, (steps, n))
    env = base._env(n_envs=n, auto_reset=False)
    env.seed(seeds=seeds)
    env.reset(params=p)
    a_dev = torch.as_tensor(acts, device=env.device)
    rew, cnt = [], []
    for t in range(steps):
        obs, r, d, info = env.step(a_dev[t])
        rew.append(r.clone()); cnt.append(info["counts"].clone())
    env.check()
    ref = oracle.batch_run(p["bw"], p["lat"], p["queue"], p["loss"], p["start_rate"], seeds, steps, actions=acts,
                           n_threads=os.cpu_count() or 1, trajectories=True)
    assert np.array_equal(torch.stack(cnt).cpu().numpy(), ref["count_traj"])
    assert np.array_equal(torch.stack(rew).cpu().numpy(), ref["reward_traj"])
    assert np.array_equal(obs.cpu().numpy(), ref["obs"])
