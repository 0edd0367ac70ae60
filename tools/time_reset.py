"""Experiment helper: where does an auto-reset step spend its time?  (python tools/time_reset.py [n_envs])"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pcc_rl_b200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
env = pcc_rl_b200.PccBatchEnv(n_envs=n, seed=100)
env.reset()
a = torch.zeros(n, dtype=torch.float64, device=env.device)
for _ in range(5):
    env.step(a)
torch.cuda.synchronize()
def ev(): return torch.cuda.Event(enable_timing=True)
# full reset through the Python API
s, e = ev(), ev()
t0 = time.perf_counter(); s.record(); env.reset(); e.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
print("env.reset(): device %.3f ms, wall %.3f ms" % (s.elapsed_time(e), 1e3 * (t1 - t0)))
p = env._sample(np.ones(n, dtype=bool))
t0 = time.perf_counter(); p = env._sample(np.ones(n, dtype=bool)); t1 = time.perf_counter()
print("host parameter sampling: %.3f ms" % (1e3 * (t1 - t0)))
s, e = ev(), ev()
s.record(); env.step(a); e.record(); torch.cuda.synchronize()
print("first step after reset: %.3f ms" % s.elapsed_time(e))
s, e = ev(), ev()
s.record(); env.step(a); e.record(); torch.cuda.synchronize()
print("second step after reset: %.3f ms" % s.elapsed_time(e))
for k in range(3):
    s, e = ev(), ev()
    s.record(); env.step(a); e.record(); torch.cuda.synchronize()
    print("  step: %.3f ms" % s.elapsed_time(e))
