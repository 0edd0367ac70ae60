"""debug: the ragged multi-sender case of tests/test_gpu_multi.py, step by step with progress output"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pcc_rl_b200
S = int(sys.argv[1]); n, steps = 96, 40
g = np.random.default_rng(100 + S)
p = dict(bw=g.uniform(80, 2000, n), lat=np.exp(g.uniform(np.log(0.002), np.log(0.6), n)),
         queue=g.integers(1, 60, n), loss=g.choice([0.0, 0.01, 0.05], n))
rates = g.uniform(40, 1500, (n, S))
acts = g.normal(0, 2.0, (steps, n, S))
env = pcc_rl_b200.PccMultiSenderEnv(n, n_senders=S, seed=900, ring_capacity=1 << 14)
env.reset(p, rates)
torch.cuda.synchronize(); print("reset done", flush=True)
for t in range(steps):
    t0 = time.time()
    obs, rew, done, info = env.step(acts[t])
    torch.cuda.synchronize()
    print(t, "%.1f ms" % (1e3 * (time.time() - t0)), int(info["counts"][:, :, 0].sum()), flush=True)
env.check()
print("ok")
