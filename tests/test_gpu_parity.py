"""-m gpu: the CUDA path (through the C ABI, libpcc_b200.so) against the CPU oracle and against
the committed outputs of the unmodified reference.  Bit-exact on integer packet counts AND on
every float output (obs, reward, MI metrics, clocks) -- stricter than the 1e-5 relative
tolerance BASELINE.json allows for the float features."""
import numpy as np
import pytest

import oracle
from golden_util import assert_step_equal, golden_names, load_golden

pytestmark = pytest.mark.gpu


def _env(**kw):
    import pcc_rl_b200
    return pcc_rl_b200.PccBatchEnv(**kw)


def default_params(n, seed):
    g = np.random.default_rng(seed)
    bw = g.uniform(100, 500, n)
    return dict(bw=bw, lat=g.uniform(0.05, 0.5, n), queue=1 + np.exp(g.uniform(0, 8, n)).astype(np.int64),
                loss=g.uniform(0, 0.05, n), start_rate=g.uniform(0.3, 1.5, n) * bw)


@pytest.mark.parametrize("name", golden_names("philox_"))
def test_cuda_matches_reference_golden_philox(name):
    g = load_golden(name)
    env = _env(n_envs=1, history_len=g["history_len"], features=g["features"], rng="philox",
               auto_reset=False, want_info=True)
    env.seed(seeds=np.array([g["seed"]], dtype=np.uint64))
    k = 0
    for ep in range(len(g["ep_params"])):
        bw, lat, q, loss, rate = g["ep_params"][ep]
        obs0 = env.reset(params=dict(bw=[bw], lat=[lat], queue=[int(q)], loss=[loss], start_rate=[rate]))
        assert np.array_equal(obs0.cpu().numpy()[0], g["ep_obs0"][ep])
        assert env.column("cur_time").item() == g["ep_cur_time0"][ep]
        for _ in range(g["steps_per_episode"]):
            obs, r, d, info = env.step(np.array([g["action"][k]]))
            m = info["metrics"].cpu().numpy()[0]
            assert_step_equal(g, k, obs.cpu().numpy()[0], r.item(), d.item(), info["counts"].cpu().numpy()[0],
                              cur_time=m[8], run_dur=m[10], rate=m[9], info=m, what=name)
            k += 1
    env.check()


def test_cuda_config2_full_episode_vs_oracle():
    """BASELINE config 2: 4 096 envs, default (ICML'19) parameter ranges, history 10, one full
    400-step episode with N(0,1) actions -- every reward and every packet count of every env-step."""
    import torch
    n, steps = 4096, 400
    p = default_params(n, 11)
    seeds = np.arange(n, dtype=np.uint64) + np.uint64(777)
    acts = np.random.default_rng(12).normal(0, 1, (steps, n))
    env = _env(n_envs=n, auto_reset=False)
    env.seed(seeds=seeds)
    env.reset(params=p)
    rew = torch.empty((steps, n), dtype=torch.float64, device=env.device)
    cnt = torch.empty((steps, n, 3), dtype=torch.int32, device=env.device)
    a_dev = torch.as_tensor(acts, device=env.device)
    for t in range(steps):
        obs, r, d, info = env.step(a_dev[t])
        rew[t] = r
        cnt[t] = info["counts"]
    env.check()
    assert bool(d.all())
    import os
    ref = oracle.batch_run(p["bw"], p["lat"], p["queue"], p["loss"], p["start_rate"], seeds, steps, actions=acts,
                           n_threads=os.cpu_count() or 1, trajectories=True)
    assert np.array_equal(cnt.cpu().numpy(), ref["count_traj"])
    assert np.array_equal(rew.cpu().numpy(), ref["reward_traj"])
    assert np.array_equal(obs.cpu().numpy(), ref["obs"])
    # reward_sum accumulates sequentially on both sides
    acc = np.zeros(n)
    for t in range(steps):
        acc += ref["reward_traj"][t]
    assert np.array_equal(env.column("last_episode_return").cpu().numpy(), acc)


def test_cuda_auto_reset_multi_episode_vs_oracle():
    """64 envs through 2.5 episodes with auto-reset (fresh link parameters per episode), all 12
    features, per-step comparison of everything."""
    feats = ",".join(oracle.METRIC_NAMES)
    n, steps = 64, 1000
    env = _env(n_envs=n, features=feats, seed=99, want_info=True)
    obs = env.reset()
    orcs = []
    for i in range(n):
        o = oracle.OracleEnv(10, feats)
        o.seed_philox(99 + i)
        o0 = o.reset(*[env.params[k][i] for k in ("bw", "lat", "queue", "loss", "start_rate")])
        assert np.array_equal(o0, obs[i].cpu().numpy())
        orcs.append(o)
    g = np.random.default_rng(3)
    for t in range(steps):
        a = g.normal(0, 2.0, n)
        obs, r, d, info = env.step(a)
        obs_h, r_h, d_h = obs.cpu().numpy(), r.cpu().numpy(), d.cpu().numpy()
        c_h, m_h = info["counts"].cpu().numpy(), info["metrics"].cpu().numpy()
        for i in range(n):
            o = orcs[i]
            o_obs, o_r, o_d, o_c, o_info = o.step(a[i])
            assert tuple(o_c) == tuple(c_h[i]), (t, i)
            assert o_r == r_h[i] and o_d == d_h[i], (t, i)
            assert np.array_equal(o_info, m_h[i][:8]), (t, i)
            if o_d:   # auto-reset: obs is the first obs of the next episode with the new parameters
                o_obs = o.reset(*[env.params[k][i] for k in ("bw", "lat", "queue", "loss", "start_rate")])
            assert np.array_equal(o_obs, obs_h[i]), (t, i)
    env.check()


def test_cuda_mt19937_batch_vs_oracle():
    n, steps = 32, 120
    p = default_params(n, 5)
    seeds = np.array([1234, 2**32 + 7, 0, 2**63] + list(range(100, 100 + n - 4)), dtype=np.uint64)
    env = _env(n_envs=n, rng="mt19937", auto_reset=False)
    env.seed(seeds=seeds)
    obs = env.reset(params=p)
    g = np.random.default_rng(8)
    orcs = []
    for i in range(n):
        o = oracle.OracleEnv()
        o.seed_mt(int(seeds[i]))
        assert np.array_equal(o.reset(*[p[k][i] for k in ("bw", "lat", "queue", "loss", "start_rate")]),
                              obs[i].cpu().numpy())
        orcs.append(o)
    for t in range(steps):
        a = g.normal(0, 1.0, n)
        obs, r, d, info = env.step(a)
        obs_h, r_h, c_h = obs.cpu().numpy(), r.cpu().numpy(), info["counts"].cpu().numpy()
        for i in range(n):
            o_obs, o_r, o_d, o_c, _ = orcs[i].step(a[i])
            assert tuple(o_c) == tuple(c_h[i]) and o_r == r_h[i] and np.array_equal(o_obs, obs_h[i]), (t, i)


def test_cuda_config3_spot_check_and_invariants():
    """BASELINE config 3 scale: 65 536 envs.  1 % of the envs are checked against the oracle;
    for all envs: conservation sent = acked + lost + change of in-flight, counts non-negative,
    obs finite, done exactly at max_steps."""
    import torch
    n, steps = 65536, 60
    env = _env(n_envs=n, seed=4242, max_steps=50)
    obs = env.reset()
    params0 = {k: v.copy() for k, v in env.params.items()}
    sel = np.random.default_rng(1).choice(n, n // 100, replace=False)
    orcs = {}
    for i in sel:
        o = oracle.OracleEnv()
        o.set_max_steps(50)
        o.seed_philox(4242 + int(i))
        o.reset(*[params0[k][i] for k in ("bw", "lat", "queue", "loss", "start_rate")])
        orcs[int(i)] = o
    g = torch.Generator(device=env.device)
    g.manual_seed(5)
    tot = torch.zeros((n, 3), dtype=torch.int64, device=env.device)
    for t in range(steps):
        a = torch.randn(n, generator=g, device=env.device, dtype=torch.float64)
        obs, r, d, info = env.step(a)
        assert bool(torch.isfinite(obs).all()) and bool(torch.isfinite(r).all())
        assert bool((info["counts"] >= 0).all())
        assert bool(d.all()) == (t == 49) and bool(d.any()) == (t == 49)
        tot += info["counts"]
        a_h, c_h, r_h, o_h = a[sel].cpu().numpy(), info["counts"][sel].cpu().numpy(), r[sel].cpu().numpy(), obs[sel].cpu().numpy()
        for j, i in enumerate(sel):
            o = orcs[int(i)]
            o_obs, o_r, o_d, o_c, _ = o.step(a_h[j])
            if o_d:
                o_obs = o.reset(*[env.params[k][i] for k in ("bw", "lat", "queue", "loss", "start_rate")])
            assert tuple(o_c) == tuple(c_h[j]) and o_r == r_h[j] and np.array_equal(o_obs, o_h[j]), (t, int(i))
        if t == 48:  # just before the synchronized reset: packets are conserved (every acked/lost
            # packet was sent in this episode, possibly during its two warm-up MIs of 3*lat each)
            warm = torch.as_tensor(2 * 3 * params0["lat"] * params0["start_rate"] + 4, device=env.device)
            assert bool((tot[:, 0] + warm >= tot[:, 1] + tot[:, 2]).all())
            assert bool((tot[:, 0] > 0).all())
    env.check()


def test_cuda_sharding_equals_single_batch():
    """§8e: an env batch split over ranks (here: two handles on one GPU with global offsets) is
    bit-identical to the same batch in one handle."""
    n, steps = 2048, 40
    whole = _env(n_envs=n, seed=31, n_global=n)
    a = _env(n_envs=n // 2, seed=31, global_offset=0, n_global=n)
    b = _env(n_envs=n // 2, seed=31, global_offset=n // 2, n_global=n)
    ow, oa, ob = whole.reset(), a.reset(), b.reset()
    import torch
    assert torch.equal(ow, torch.cat([oa, ob]))
    g = np.random.default_rng(2)
    for t in range(steps):
        act = g.normal(0, 1, n)
        xw = whole.step(act)
        xa = a.step(act[: n // 2])
        xb = b.step(act[n // 2:])
        assert torch.equal(xw[0], torch.cat([xa[0], xb[0]]))
        assert torch.equal(xw[1], torch.cat([xa[1], xb[1]]))
        assert torch.equal(xw[3]["counts"], torch.cat([xa[3]["counts"], xb[3]["counts"]]))


def test_cuda_step_host_equals_device_path():
    import torch
    n = 512
    e1 = _env(n_envs=n, seed=8, auto_reset=False)
    e2 = _env(n_envs=n, seed=8, auto_reset=False)
    e1.reset()
    e2.reset()
    acts = torch.zeros(n, dtype=torch.float64).pin_memory()
    obs = torch.zeros((n, 30), dtype=torch.float64).pin_memory()
    rew = torch.zeros(n, dtype=torch.float64).pin_memory()
    done = torch.zeros(n, dtype=torch.uint8).pin_memory()
    cnt = torch.zeros((n, 3), dtype=torch.int32).pin_memory()
    g = np.random.default_rng(4)
    for t in range(20):
        a = g.normal(0, 1, n)
        acts.numpy()[:] = a
        e1.step_host(acts.numpy(), obs.numpy(), rew.numpy(), done.numpy(), cnt.numpy())
        o2, r2, d2, i2 = e2.step(a)
        assert np.array_equal(obs.numpy(), o2.cpu().numpy()) and np.array_equal(rew.numpy(), r2.cpu().numpy())
        assert np.array_equal(cnt.numpy(), i2["counts"].cpu().numpy())


def test_cuda_checkpoint_attach_resumes_identically():
    import ctypes as C
    import torch
    from pcc_rl_b200 import _lib
    n = 256
    env = _env(n_envs=n, seed=17, auto_reset=False)
    env.reset()
    g = np.random.default_rng(6)
    for t in range(13):
        env.step(g.normal(0, 1, n))
    torch.cuda.synchronize()
    state, ring = env.state_ws.clone(), env.ring_ws.clone()   # the checkpoint
    h2 = C.c_void_p()
    _lib.check(env.L.pcc_attach(C.byref(h2), C.byref(env.cfg), state.data_ptr(), ring.data_ptr()))
    obs2 = torch.empty_like(env.obs); rew2 = torch.empty_like(env.reward)
    done2 = torch.empty_like(env.done); cnt2 = torch.empty_like(env.counts)
    for t in range(10):
        a = torch.as_tensor(g.normal(0, 1, n), device=env.device)
        env.step_device(a)
        _lib.check(env.L.pcc_step(h2, a.data_ptr(), obs2.data_ptr(), rew2.data_ptr(), done2.data_ptr(),
                                  cnt2.data_ptr(), None, None))
        torch.cuda.synchronize()
        assert torch.equal(env.obs, obs2) and torch.equal(env.reward, rew2) and torch.equal(env.counts, cnt2)
    env.L.pcc_destroy(h2)


def test_cuda_ring_overflow_is_reported():
    from pcc_rl_b200 import _lib
    env = _env(n_envs=8, ring_capacity=64, auto_reset=False)
    env.reset(params=dict(bw=[500.0] * 8, lat=[0.5] * 8, queue=[1000] * 8, loss=[0.0] * 8, start_rate=[750.0] * 8))
    with pytest.raises(_lib.PccError) as ei:
        env.check()
    assert ei.value.code == _lib.PCC_EOVERFLOW


def test_cuda_rollout_equals_stepwise():
    """pcc_rollout (K monitor intervals in one launch, in-kernel auto-reset) is bit-identical to K calls
    of pcc_step + pcc_reset, including across an episode boundary."""
    import torch
    n, K1, K2 = 512, 390, 60          # the second rollout crosses step 400: every env resets inside it
    acts = np.random.default_rng(21).normal(0, 1.5, (K1 + K2, n))
    a = _env(n_envs=n, seed=77, want_info=False)
    b = _env(n_envs=n, seed=77, want_info=False)
    assert torch.equal(a.reset(), b.reset())
    k = 0
    for K in (K1, K2):
        out = a.rollout(K, actions=torch.as_tensor(acts[k:k + K], device=a.device))
        for t in range(K):
            obs, r, d, info = b.step(acts[k + t])
            assert torch.equal(out["counts"][t], info["counts"]), (k, t)
            assert torch.equal(out["reward"][t], r) and torch.equal(out["done"][t], d), (k, t)
            assert torch.equal(out["obs"][t], obs), (k, t)
        k += K
    assert bool(out["done"].any())
    assert np.array_equal(a._steps, b._steps) and np.array_equal(a._episode, b._episode)
    for name in ("cur_time", "run_dur", "rate", "episode_return", "last_episode_return", "bw"):
        assert torch.equal(a.column(name), b.column(name)), name
    # and the two envs stay interchangeable afterwards
    o1 = a.step(acts[0])
    o2 = b.step(acts[0])
    assert torch.equal(o1[0], o2[0]) and torch.equal(o1[1], o2[1])
    a.check()


def test_cuda_rollout_with_on_device_policy():
    """Closed loop: the MLP policy (stable_solve.py:30-45: 30 -> 32 -> 16 -> 1, tanh) runs inside the rollout
    kernel.  Its actions must be the torch fp64 MLP of the previous observation, and replaying those actions
    step by step must reproduce the rollout exactly."""
    import torch
    n, K = 256, 80
    g = torch.Generator().manual_seed(3)
    pol = dict(w1=torch.randn(32, 30, generator=g, dtype=torch.float64) * 0.3, b1=torch.randn(32, generator=g, dtype=torch.float64) * 0.1,
               w2=torch.randn(16, 32, generator=g, dtype=torch.float64) * 0.3, b2=torch.randn(16, generator=g, dtype=torch.float64) * 0.1,
               w3=torch.randn(1, 16, generator=g, dtype=torch.float64) * 2.0, b3=torch.randn(1, generator=g, dtype=torch.float64) * 0.1)
    a = _env(n_envs=n, seed=5, max_steps=50)     # short episodes: resets happen inside the rollout
    b = _env(n_envs=n, seed=5, max_steps=50)
    obs0 = a.reset().clone()
    b.reset()
    out = a.rollout(K, policy=pol)
    dev = a.device
    W = {k: v.to(dev) for k, v in pol.items()}
    mlp = lambda o: (torch.tanh(torch.tanh(o @ W["w1"].T + W["b1"]) @ W["w2"].T + W["b2"]) @ W["w3"].T + W["b3"]).squeeze(-1)
    prev = obs0
    for t in range(K):
        want = mlp(prev)
        assert torch.allclose(out["actions"][t], want, rtol=1e-12, atol=1e-12), t
        obs, r, d, info = b.step(out["actions"][t])
        assert torch.equal(out["obs"][t], obs) and torch.equal(out["reward"][t], r) and torch.equal(out["done"][t], d), t
        prev = out["obs"][t]
    assert bool(out["done"].any())
    # stochastic actions: Gaussian noise around the same mean, reproducible for a given noise seed
    c = _env(n_envs=n, seed=5, max_steps=50); c.reset()
    d_ = _env(n_envs=n, seed=5, max_steps=50); d_.reset()
    sp = dict(pol, stochastic=True, log_std=-1.0, noise_seed=9)
    o1 = c.rollout(10, policy=sp)
    o2 = d_.rollout(10, policy=sp)
    assert torch.equal(o1["actions"], o2["actions"]) and torch.equal(o1["reward"], o2["reward"])
    noise = (o1["actions"][0] - mlp(obs0)) / np.exp(-1.0)
    assert 0.8 < float(noise.std()) < 1.2 and abs(float(noise.mean())) < 0.25
