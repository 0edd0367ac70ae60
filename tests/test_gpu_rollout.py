"""-m gpu: pcc_rollout (policy / value kernels + env step + auto-reset enqueued per step on the device) against the
CPU oracle -- not only against the stepwise CUDA path (tests/test_gpu_parity.py) -- and its value head against a
torch fp64 MLP.  Runs through both execution modes of the step kernel."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


def _env(**kw):
    import pcc_rl_b200
    return pcc_rl_b200.PccBatchEnv(**kw)


@pytest.mark.parametrize("mode", ["warp", "packed"])
def test_rollout_vs_oracle_across_episode_boundary(mode, monkeypatch):
    """96 envs x 450 steps in ONE pcc_rollout call (every env finishes an episode at step 400 and is reset in the
    call from the parameter bank): rewards, packet counts, done flags and observations of every step == oracle."""
    import torch
    monkeypatch.setenv("PCC_B200_MODE", mode)
    monkeypatch.setenv("PCC_B200_QUAD", "100")
    monkeypatch.setenv("PCC_B200_SOLO", "600")
    n, K = 96, 450
    env = _env(n_envs=n, seed=321)
    obs0 = env.reset()
    names = ("bw", "lat", "queue", "loss", "start_rate")
    orcs = []
    for i in range(n):
        o = oracle.OracleEnv(10, oracle.DEFAULT_FEATURES)
        o.seed_philox(321 + i)
        assert np.array_equal(o.reset(*[env.params[k][i] for k in names]), obs0[i].cpu().numpy())
        orcs.append(o)
    acts = np.random.default_rng(4).normal(0, 1.5, (K, n))
    out = env.rollout(K, actions=torch.as_tensor(acts, device=env.device))
    env.check()
    r, d, c, ob = (out[k].cpu().numpy() for k in ("reward", "done", "counts", "obs"))
    assert d[399].all() and not d[:399].any() and not d[400:].any()
    for t in range(K):
        for i, o in enumerate(orcs):
            o_obs, o_r, o_d, o_c, _ = o.step(acts[t, i])
            assert tuple(o_c) == tuple(c[t, i]), (t, i)
            assert o_r == r[t, i] and bool(o_d) == bool(d[t, i]), (t, i)
            if o_d:   # the observation of a finished env is the first one of its next episode (new parameters)
                o_obs = o.reset(*[env.params[k][i] for k in names])
            assert np.array_equal(o_obs, ob[t, i]), (t, i)


def test_rollout_value_head_and_policy_vs_torch():
    """vpred[k] = V(observation before step k), vpred[K] = V(last observation): the on-device value network against
    a torch fp64 MLP; the actions stay those of the policy network."""
    import torch
    n, K = 200, 40
    g = torch.Generator().manual_seed(11)
    rn = lambda *s, sc: torch.randn(*s, generator=g, dtype=torch.float64) * sc
    pol = dict(w1=rn(32, 30, sc=0.3), b1=rn(32, sc=0.1), w2=rn(16, 32, sc=0.3), b2=rn(16, sc=0.1), w3=rn(1, 16, sc=2.0), b3=rn(1, sc=0.1),
               vw1=rn(32, 30, sc=0.4), vb1=rn(32, sc=0.2), vw2=rn(16, 32, sc=0.4), vb2=rn(16, sc=0.2), vw3=rn(1, 16, sc=1.0), vb3=rn(1, sc=0.3))
    env = _env(n_envs=n, seed=8, max_steps=25)        # resets inside the rollout
    obs0 = env.reset().clone()
    out = env.rollout(K, policy=pol)
    env.check()
    W = {k: v.to(env.device) for k, v in pol.items()}
    mlp = lambda o, p: (torch.tanh(torch.tanh(o @ W[p + "w1"].T + W[p + "b1"]) @ W[p + "w2"].T + W[p + "b2"]) @ W[p + "w3"].T
                        + W[p + "b3"]).squeeze(-1)
    assert out["vpred"].shape == (K + 1, n)
    prev = obs0
    for t in range(K):
        assert torch.allclose(out["actions"][t], mlp(prev, ""), rtol=1e-12, atol=1e-12), t
        assert torch.allclose(out["vpred"][t], mlp(prev, "v"), rtol=1e-12, atol=1e-12), t
        prev = out["obs"][t]
    assert torch.allclose(out["vpred"][K], mlp(prev, "v"), rtol=1e-12, atol=1e-12)
    assert bool(out["done"].any())
