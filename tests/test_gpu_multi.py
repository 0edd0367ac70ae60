"""-m gpu: BASELINE config 5 -- link-parameter grid sweep (bw 1-1000 Mbit/s x delay 1-500 ms), two senders per
link -- the CUDA multi-sender path against the reference's golden outputs and against the oracle."""
import os

import numpy as np
import pytest

import oracle
from golden_util import GOLDEN_DIR, golden_names

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["warp", "thread", "heap"])
def multi_mode(request, monkeypatch):
    """The engines of the multi-sender path: the heap-free streaming MI with one link per warp (default) or per thread,
    and the per-env event heap."""
    monkeypatch.setenv("PCC_MULTI_MODE", request.param)
    return request.param


@pytest.mark.parametrize("name", golden_names("multi_"))
def test_cuda_multi_sender_matches_reference_golden(name, multi_mode):
    import pcc_rl_b200
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    bw, lat, queue, loss = z["params"]
    S = len(z["rates"])
    env = pcc_rl_b200.PccMultiSenderEnv(1, n_senders=S, ring_capacity=1 << 15)
    env.seed(seeds=np.array([int(z["seed"])], dtype=np.uint64))
    env.reset(dict(bw=[bw], lat=[lat], queue=[int(queue)], loss=[loss]), z["rates"][None, :])
    for t in range(len(z["action"])):
        obs, rew, done, info = env.step(z["action"][t][None, :])
        assert np.array_equal(info["counts"].cpu().numpy()[0], z["counts"][t]), (name, t)
        assert np.array_equal(obs.cpu().numpy()[0], z["obs"][t]) and np.array_equal(rew.cpu().numpy()[0], z["reward"][t])
    env.check()


_ORACLE_CACHE = {}


def _oracle_run(key, p, rates, seed0, acts):
    """Oracle trajectories (obs, reward, counts per step) of a batch, computed once and shared by the engine modes."""
    if key not in _ORACLE_CACHE:
        n = len(rates)
        orcs = []
        for i in range(n):
            o = oracle.OracleEnv()
            o.seed_philox(seed0 + i)
            o.reset_multi(p["bw"][i], p["lat"][i], int(p["queue"][i]), p["loss"][i], rates[i])
            orcs.append(o)
        out = []
        for a in acts:
            res = [orcs[i].step_multi(a[i]) for i in range(n)]
            out.append((np.stack([r[0] for r in res]), np.stack([r[1] for r in res]), np.stack([r[3] for r in res])))
        _ORACLE_CACHE[key] = out
    return _ORACLE_CACHE[key]


def _compare(env, acts, ref):
    for t, a in enumerate(acts):
        obs, rew, done, info = env.step(a)
        o_obs, o_rew, o_cnt = ref[t]
        cnt_h = info["counts"].cpu().numpy()
        bad = np.nonzero((cnt_h != o_cnt).any(axis=(1, 2)))[0]
        assert bad.size == 0, (t, bad[:8])
        assert np.array_equal(o_obs, obs.cpu().numpy()) and np.array_equal(o_rew, rew.cpu().numpy()), t
    env.check()


def test_cuda_config5_grid_sweep_vs_oracle(multi_mode):
    """32 x 32 grid of (bandwidth, delay), 2 senders per link, 60 steps: every grid point against the oracle."""
    import pcc_rl_b200
    p = pcc_rl_b200.grid_sweep_params(n_bw=32, n_lat=32, queue=40, loss=0.01)
    n, S, steps = 1024, 2, 60
    g = np.random.default_rng(7)
    rates = g.uniform(40, 1000, (n, S))
    acts = g.normal(0, 2.0, (steps, n, S))
    env = pcc_rl_b200.PccMultiSenderEnv(n, n_senders=S, seed=500, ring_capacity=1 << 13)
    env.reset(p, rates)
    _compare(env, acts, _oracle_run("grid", p, rates, 500, acts))


@pytest.mark.parametrize("S", [1, 3, 4])
def test_cuda_multi_other_sender_counts_vs_oracle(S, multi_mode):
    """1, 3 and 4 senders per link (the kernel is compiled per sender count), ragged link parameters incl. lossy,
    tiny-queue (0 and 1 packet: tail_drop_threshold's corner) and long-delay links whose MIs span several 64-draw rounds and several numpy leaves."""
    import pcc_rl_b200
    n, steps = 96, 40
    g = np.random.default_rng(100 + S)
    p = dict(bw=g.uniform(80, 2000, n), lat=np.exp(g.uniform(np.log(0.002), np.log(0.6), n)),
             queue=g.integers(0, 60, n), loss=g.choice([0.0, 0.01, 0.05, 1.0], n))
    rates = g.uniform(40, 1500, (n, S))
    acts = g.normal(0, 2.0, (steps, n, S))
    env = pcc_rl_b200.PccMultiSenderEnv(n, n_senders=S, seed=900, ring_capacity=1 << 14)
    env.reset(p, rates)
    _compare(env, acts, _oracle_run(("ragged", S), p, rates, 900, acts))
