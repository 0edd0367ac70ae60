"""-m gpu: the congestion-window / latency-noise variants of the event loop (USE_CWND, USE_LATENCY_NOISE,
network_sim.py:51-54) on the CUDA heap path, against outputs of the unmodified reference with
the switches on (tests/golden/variant_*.npz) and against the oracle on a batch."""
import os

import numpy as np
import pytest

import oracle
from golden_util import GOLDEN_DIR, golden_names

pytestmark = pytest.mark.gpu

SINGLE = [n for n in golden_names("variant_") if not n.startswith("variant_multi_")]


@pytest.mark.parametrize("name", SINGLE)
def test_cuda_variant_single_sender_golden(name):
    """The reference's own SimulatedNetworkEnv with a switch turned on = one sender on the heap path."""
    import pcc_rl_b200
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    use_cwnd, use_noise = bool(z["use_cwnd"]), bool(z["use_noise"])
    env = pcc_rl_b200.PccMultiSenderEnv(1, n_senders=1, features=str(z["features"]), ring_capacity=1 << 15,
                                        use_cwnd=use_cwnd, use_latency_noise=use_noise)
    env.seed(seeds=np.array([int(z["seed"])], dtype=np.uint64))
    k = 0
    for ep in range(len(z["ep_params"])):
        bw, lat, q, loss, rate = z["ep_params"][ep]
        obs0 = env.reset(dict(bw=[bw], lat=[lat], queue=[int(q)], loss=[loss]), [[rate]])   # the stream continues
        assert np.array_equal(obs0.cpu().numpy()[0, 0], z["ep_obs0"][ep])
        for _ in range(int(z["steps_per_episode"])):
            obs, rew, done, info = env.step([[z["action"][k]]], [[z["cwnd_action"][k]]] if use_cwnd else None)
            assert tuple(info["counts"].cpu().numpy()[0, 0]) == tuple(z["counts"][k]), (name, k)
            assert np.array_equal(obs.cpu().numpy()[0, 0], z["obs"][k]), (name, k)
            assert rew.cpu().numpy()[0, 0] == z["reward"][k], (name, k)
            assert int(info["cwnd"][0, 0]) == int(z["cwnd"][k]), (name, k)
            assert bool(done[0]) == bool(z["done"][k])
            k += 1
    env.check()


@pytest.mark.parametrize("name", golden_names("variant_multi_"))
def test_cuda_variant_multi_golden(name):
    import pcc_rl_b200
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    bw, lat, queue, loss = z["params"]
    S = len(z["rates"])
    env = pcc_rl_b200.PccMultiSenderEnv(1, n_senders=S, ring_capacity=1 << 15, use_cwnd=bool(z["use_cwnd"]),
                                        use_latency_noise=bool(z["use_noise"]))
    env.seed(seeds=np.array([int(z["seed"])], dtype=np.uint64))
    env.reset(dict(bw=[bw], lat=[lat], queue=[int(queue)], loss=[loss]), z["rates"][None, :])
    for t in range(len(z["action"])):
        obs, rew, done, info = env.step(z["action"][t][None, :], z["cwnd_action"][t][None, :])
        assert np.array_equal(info["counts"].cpu().numpy()[0], z["counts"][t]), (name, t)
        assert np.array_equal(obs.cpu().numpy()[0], z["obs"][t]) and np.array_equal(rew.cpu().numpy()[0], z["reward"][t])
        assert list(info["cwnd"].cpu().numpy()[0]) == list(z["cwnd"][t])
    env.check()


@pytest.mark.parametrize("use_cwnd,use_noise", [(True, False), (False, True), (True, True)])
def test_cuda_variant_batch_vs_oracle(use_cwnd, use_noise):
    """512 single-sender envs with default-range link parameters, 50 steps, every env against the oracle;
    the window action arrives as the reference's 2-dim action."""
    import pcc_rl_b200
    n, steps = 512, 50
    p = pcc_rl_b200.sample_link_params(77, 1, np.arange(n), n)
    env = pcc_rl_b200.PccMultiSenderEnv(n, n_senders=1, seed=900, ring_capacity=1 << 15, use_cwnd=use_cwnd,
                                        use_latency_noise=use_noise)
    env.reset(p, p["start_rate"][:, None])
    orcs = []
    for i in range(n):
        o = oracle.OracleEnv()
        o.set_variant(use_cwnd, use_noise)
        o.seed_philox(900 + i)
        o.reset(p["bw"][i], p["lat"][i], int(p["queue"][i]), p["loss"][i], p["start_rate"][i])
        orcs.append(o)
    g = np.random.default_rng(5)
    for t in range(steps):
        a = g.normal(0, 2.0, (n, 1, 2)) * np.array([1.0, 3.0])
        obs, rew, done, info = env.step(a if use_cwnd else a[..., 0])
        obs_h, rew_h, cnt_h = obs.cpu().numpy(), rew.cpu().numpy(), info["counts"].cpu().numpy()
        cw_h = info["cwnd"].cpu().numpy()
        for i in range(n):
            if use_cwnd:
                o_obs, o_r, o_d, o_c, _ = orcs[i].step_cwnd(a[i, 0, 0], a[i, 0, 1])
            else:
                o_obs, o_r, o_d, o_c, _ = orcs[i].step(a[i, 0, 0])
            assert tuple(o_c) == tuple(cnt_h[i, 0]), (t, i)
            assert np.array_equal(o_obs, obs_h[i, 0]) and o_r == rew_h[i, 0], (t, i)
            assert orcs[i].cwnd() == cw_h[i, 0]
    env.check()
